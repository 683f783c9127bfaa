for cap in 0 48 32 64 0; do
echo "== CTX_MAX_CTAS=$cap"
CVB_CTX_MAX_CTAS=$cap timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch-obs 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"ms_per_step\"], d[\"p50_ms\"], d[\"e2e\"][\"p50_ms\"], d[\"phases_ms\"], d[\"clocks\"])"
done
