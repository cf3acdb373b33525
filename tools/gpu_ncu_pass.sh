# ncu evidence for the training pass (run only after the plain bench exited 0):
#   (1) launch list of one FOMAML task (2 passes + inner SGD step) with gpu__time_duration.sum,
#   (2) ncu --set full of the six persistent recurrent kernels of the second pass, exported as raw CSV.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_ncu_pass.sh v12'
V=${1:-vX}
MSA_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$V.csv \
    python profiles/run_pass.py 1 > gpurun_out/ncu_list_$V.log 2>&1; echo list-exit $?
python profiles/summarize_launches.py gpurun_out/launches_$V.csv > gpurun_out/launches_$V.txt 2>&1; head -12 gpurun_out/launches_$V.txt
MSA_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_attn_chain|k_lstm_rec" -s 6 -c 6 \
    -o gpurun_out/prof_chains_$V -f python profiles/run_pass.py 1 > gpurun_out/ncu_full_$V.log 2>&1; echo full-exit $?
ncu -i gpurun_out/prof_chains_$V.ncu-rep --page raw --csv > gpurun_out/prof_chains_$V.csv 2> gpurun_out/prof_chains_$V.err; wc -c gpurun_out/prof_chains_$V.csv
