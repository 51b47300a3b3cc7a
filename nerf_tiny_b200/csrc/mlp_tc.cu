// placeholder, replaced by the tcgen05 kernel
#include "common.cuh"
size_t nt_mlp_tc_packed_bytes() { return 256; }
int nt_mlp_tc_pack(nt_ctx*, const float*, void*, cudaStream_t) { nt_set_error("bf16 path not built"); return NT_ERR_UNSUPPORTED; }
int nt_mlp_tc_forward(nt_ctx*, int64_t, int, const float*, const float*, const float*, const float*, const void*, float*, float*, cudaStream_t) { nt_set_error("bf16 path not built"); return NT_ERR_UNSUPPORTED; }
