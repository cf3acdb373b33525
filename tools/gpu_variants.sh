# A/B timing of differently compiled builds of libmsa_b200.so (variants/lib_<tag>.so, built on the CPU box with MSA_NVCC_DEFS):
# per-kernel event times of the recurrences for single passes (tensor-core kernels forced) and, with MSA_GRPS="2 8", for groups.
# Usage: gpurun --timeout 900 -- 'MSA_GRPS="2 8" bash tools/gpu_variants.sh tagA tagB ...'
for tag in "$@"; do
  echo "== $tag"
  MSA_LIB_PATH=$PWD/variants/lib_$tag.so MSA_CHAIN_MMA=2 python profiles/group_bench.py 1 2>&1 | tail -1 | sed 's/  [a-z_]*_grp 0us//g; s/grouped pass/gp/'
  if [ -n "$MSA_GRPS" ]; then MSA_LIB_PATH=$PWD/variants/lib_$tag.so python profiles/group_bench.py $MSA_GRPS 2>&1 | tail -2 | sed 's/  [a-z_]*_fwd 0us//g; s/  [a-z_]*_bwd 0us//g; s/grouped pass/gp/'; fi
done
