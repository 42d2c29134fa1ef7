#!/bin/bash
# prints the PTX nvcc emits for log2f(x) (libdevice __nv_log2f inlined): the routine
# oracle/nmi_oracle.c:log2f_cuda transcribes operation by operation
d=$(mktemp -d); cat > $d/l.cu <<'CU'
__global__ void k(float* o, const float* i) { o[threadIdx.x] = log2f(i[threadIdx.x]); }
CU
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -ptx $d/l.cu -o $d/l.ptx && sed -n '/ld.global.f32/,/st.global.f32/p' $d/l.ptx; rm -rf $d
