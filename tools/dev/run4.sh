for d in 0 1 2; do
echo "dbg=$d"
timeout 120 python tools/trace_attn.py 72b-tp4 2 attn_dbg=$d | head -17
done
