V=nerf_tiny_b200/build/variants
for rep in 1 2; do for v in "" dwnohint; do
  if [ -z "$v" ]; then L=""; else L=$PWD/$V/lib$v.so; fi
  for n in 1024 4096; do
  echo "lib=${v:-default} n=$n: $(NT_LIB_PATH=$L NT_DW_OVERLAP_CTAS=0 timeout 100 python tools/train_kernel_times.py $n 2>&1 | grep -E "graph|dw_grouped|Error" | cut -c1-70 | tr "\n" " ")"
done; done; done
