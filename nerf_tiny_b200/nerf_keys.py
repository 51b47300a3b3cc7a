"""state_dict keys of the reference network in creation order (nerf.py:85-99)."""
LAYER_KEYS = [
    "network.point_layer.0.0", "network.point_layer.1.0", "network.point_layer.2.0", "network.point_layer.3.0",
    "network.point_layer.4.0", "network.point_layer.5.0", "network.point_layer.6.0", "network.point_layer.7.0",
    "network.sigma_layer.0", "network.point_info", "network.dir_info.0", "network.color_layer.0",
]
LAYER_SHAPES = [(256, 60), (256, 256), (256, 256), (256, 256), (256, 316), (256, 256), (256, 256), (256, 256),
                (1, 256), (256, 256), (128, 280), (3, 128)]
