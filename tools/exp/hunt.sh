nvidia-smi --query-gpu=serial --format=csv,noheader
N=${N:-40}
n=0; for i in $(seq 1 $N); do timeout 100 python -m pytest tools/exp/test_hunt.py -q -p no:cacheprovider > /tmp/h.log 2>&1; if grep -q "1 failed" /tmp/h.log; then n=$((n+1)); grep -E "^E  " /tmp/h.log | cut -c1-200 | head -4; fi; done; echo "default build: $n of $N bad"
