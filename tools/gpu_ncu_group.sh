# ncu evidence for the grouped tensor-core recurrences (run only after the same plain command exited 0 in this call):
#   (1) launch list of one 8-task FOMAML meta-step (1 grouped train pass + 8 test passes) with gpu__time_duration.sum,
#   (2) ncu --set full of the grouped kernels of the train pass (k_*_mma), exported as raw CSV.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_ncu_group.sh v1'
V=${1:-vX}
MSA_REPS=1 timeout 300 python profiles/run_pass.py 8 > gpurun_out/plain_$V.log 2>&1 || { echo plain-run-failed; tail -5 gpurun_out/plain_$V.log; exit 1; }
MSA_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_$V.csv \
    python profiles/run_pass.py 8 > gpurun_out/ncu_list_$V.log 2>&1; echo list-exit $?
python profiles/summarize_launches.py gpurun_out/r02_launches_$V.csv > gpurun_out/r02_launches_$V.txt 2>&1; head -14 gpurun_out/r02_launches_$V.txt
MSA_REPS=1 timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"_mma" -c 6 \
    -o gpurun_out/r02_prof_mma_$V -f python profiles/run_pass.py 8 > gpurun_out/ncu_full_$V.log 2>&1; echo full-exit $?
ncu -i gpurun_out/r02_prof_mma_$V.ncu-rep --page raw --csv > gpurun_out/r02_prof_mma_$V.csv 2> gpurun_out/r02_prof_mma_$V.err; wc -c gpurun_out/r02_prof_mma_$V.csv
