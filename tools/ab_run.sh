V=nerf_tiny_b200/build/variants
for rep in 1 2 3; do for v in "" smembias; do
  if [ -z "$v" ]; then python tools/mlp_ab.py 160000 10; else NT_LIB_PATH=$PWD/$V/lib$v.so python tools/mlp_ab.py 160000 10; fi
done; done
