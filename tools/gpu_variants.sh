# A/B timing of differently compiled builds of libmsa_b200.so (variants/lib_<tag>.so, built on the CPU box with MSA_NVCC_DEFS):
# per-kernel event times of the recurrences for single passes with the tensor-core kernels forced and for groups of 2 and 8.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_variants.sh tagA tagB ...'
for tag in "$@"; do
  echo "== $tag"
  MSA_LIB_PATH=$PWD/variants/lib_$tag.so MSA_CHAIN_MMA=2 python profiles/group_bench.py 1 2>&1 | tail -1 | sed 's/  [a-z_]*_grp 0us//g; s/grouped pass/gp/'
  MSA_LIB_PATH=$PWD/variants/lib_$tag.so python profiles/group_bench.py 2 8 2>&1 | tail -2 | sed 's/  [a-z_]*_fwd 0us//g; s/  [a-z_]*_bwd 0us//g; s/grouped pass/gp/'
done
