for bo in 100 400; do
  MSA_POLL_BACKOFF=$bo python metaspeakeradaptation-tts_b200/build.py --force > /dev/null 2>&1
  echo "backoff $bo"; timeout 300 python profiles/chain_vs_batch.py 4 2>&1 | grep "^B="
done
python metaspeakeradaptation-tts_b200/build.py --force > /dev/null 2>&1
echo "no backoff"; timeout 300 python profiles/chain_vs_batch.py 4 2>&1 | grep "^B="
