V=nerf_tiny_b200/build/variants
for rep in 1 2; do
for v in "" base nopf epispin epispin_nopf; do
  if [ -z "$v" ]; then python tools/mlp_ab.py; else NT_LIB_PATH=$PWD/$V/lib$v.so python tools/mlp_ab.py; fi
done; done
