# Debug build with per-phase clocks in the leaf kernel (tools/dbg_clocks.py): tools/libgegp_dbg.so
set -e
mkdir -p /tmp/dbg
for f in gemm gemm_tma leaf potrf build lml eig peak capi; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    -Wno-deprecated-gpu-targets -DGEGP_LEAF_CLOCKS -c gpgradpy_b200/csrc/$f.cu -o /tmp/dbg/$f.o 2>/dev/null &
done
wait
/usr/local/cuda/bin/nvcc -shared -Wno-deprecated-gpu-targets -o tools/libgegp_dbg.so /tmp/dbg/*.o
