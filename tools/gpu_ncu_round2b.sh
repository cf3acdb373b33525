# ncu evidence for the second half of round 2 (run only after the same plain command exited 0 in this call):
#   (1) launch list of one 8-task FOMAML meta-step with gpu__time_duration.sum,
#   (2) ncu --set full of the per-task-weight attention chain (k_attn_fwd_mma<3, true>) and of the implicit-convolution launches of
#       the tcgen05 GEMM kernel (forward 3xTF32, weight gradient / input gradient single TF32), exported as raw CSV.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_ncu_round2b.sh v2'
V=${1:-vX}
MSA_REPS=1 timeout 300 python profiles/run_pass.py 8 > gpurun_out/plain_$V.log 2>&1 || { echo plain-run-failed; tail -5 gpurun_out/plain_$V.log; exit 1; }
MSA_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_$V.csv \
    python profiles/run_pass.py 8 > gpurun_out/ncu_list_$V.log 2>&1; echo list-exit $?
python profiles/summarize_launches.py gpurun_out/r02_launches_$V.csv > gpurun_out/r02_launches_$V.txt 2>&1; head -14 gpurun_out/r02_launches_$V.txt
# the PT forward chain: the launches of k_attn_fwd_mma after the shared-weight one (skip 1, take 1)
MSA_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_attn_fwd_mma" -s 1 -c 1 \
    -o gpurun_out/r02_prof_pt_$V -f python profiles/run_pass.py 8 > gpurun_out/ncu_full_pt_$V.log 2>&1; echo full-pt-exit $?
ncu -i gpurun_out/r02_prof_pt_$V.ncu-rep --page raw --csv > gpurun_out/r02_prof_pt_$V.csv 2> gpurun_out/r02_prof_pt_$V.err; wc -c gpurun_out/r02_prof_pt_$V.csv
# conv launches of the GEMM kernel in one single-task pass: forward (3xTF32) of a postnet layer, and its two backward products
MSA_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_gemm_tf32_nt" -s 14 -c 8 \
    -o gpurun_out/r02_prof_conv_$V -f python profiles/run_pass.py 1 > gpurun_out/ncu_full_conv_$V.log 2>&1; echo full-conv-exit $?
ncu -i gpurun_out/r02_prof_conv_$V.ncu-rep --page raw --csv > gpurun_out/r02_prof_conv_$V.csv 2> gpurun_out/r02_prof_conv_$V.err; wc -c gpurun_out/r02_prof_conv_$V.csv
