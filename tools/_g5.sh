for cfg in "fly=2" "fly=2 12=1 13=1" "fly=2 13=1" "fly=2 12=1 13=1 11=250" "fly=2 12=1 13=1 11=310" "fly=1 12=1 13=1"; do
  timeout 300 python tools/shard_times.py 20 8 $cfg 2>&1 | tail -9
done
