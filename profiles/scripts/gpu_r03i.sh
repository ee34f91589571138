# r03i: final record of the round on one GPU: full GPU suite, the driver's default bench, the reference arm, smoke()
O=gpurun_out; T=${1:-r03i}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
grep -E "passed|failed|EXIT" $O/${T}_gpu_tests.log | tail -4
s0=$(date +%s); timeout 1500 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench wall $(( $(date +%s) - s0 )) s" | tee -a $O/${T}_bench.err
s0=$(date +%s); timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.log 2> $O/${T}_bench_reference.err; echo "reference arm wall $(( $(date +%s) - s0 )) s" | tee -a $O/${T}_bench_reference.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo EXIT=$? >> $O/${T}_smoke.log; tail -2 $O/${T}_smoke.log
python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print(round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], d["step_breakdown_ms"], d["parity"]["pass"], d["clocks"])
print(json.dumps(d["index_build"])[:500]); print(json.dumps(d["rerank"])[:250]); print(json.dumps(d["vit_g"])[:400])
l=[x for x in open("$O/${T}_bench_reference.log") if x.startswith("{")][-1]; print(l[:300])
PY
