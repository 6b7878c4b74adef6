cd /root/repo
for d in 0 1 3 28 2 0 1 3; do echo "== dbg $d"; IMDBN_DEBUG_STREAM=$d IMDBN_TS_TRACE=100 python tools/pass_time.py 2>&1 | grep "first stage\|last mma\|tf32" | head -8; done
