for c in 0 1 2 3; do
  echo "cfg fwd=$c bwd=$c"; TMD_COOP_FWD=$c TMD_COOP_BWD=$c timeout 120 python scripts/time_mp.py 2>&1 | tail -1
done
