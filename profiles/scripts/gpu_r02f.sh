# r02f: second-generation ViT attention: op tests, timing v1 vs v2, index-build rate
O=gpurun_out; T=${1:-r02f}; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -s -k "vit_attention" > $O/${T}_vit_attn_tests.log 2>&1; echo EXIT=$? >> $O/${T}_vit_attn_tests.log
grep -E "^\[|passed|failed|EXIT|Error|error|assert" $O/${T}_vit_attn_tests.log | tail -20
if grep -q "EXIT=0" $O/${T}_vit_attn_tests.log; then
  for dh in 64 88; do
    SPRC_VIT_ATTN_V1=1 timeout 120 python tests/gpu_prof_attn.py 128 $dh >> $O/${T}_vit_attn_timing.log 2>&1
    timeout 120 python tests/gpu_prof_attn.py 128 $dh >> $O/${T}_vit_attn_timing.log 2>&1
  done
  cat $O/${T}_vit_attn_timing.log
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s -k "stage_parity or recall_parity_full" > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
  grep -E "^\[|passed|failed|EXIT" $O/${T}_parity.log | tail -14
  timeout 600 python bench.py --no-cpu-baseline --no-vitg --no-rerank --no-eager-gpu --index-images 8192 --steps 5 --profile-dump $O/${T}_prof > $O/${T}_bench.log 2>&1
  python tools/show_profile.py $O/${T}_prof.index.csv | head -8
  python - <<PY
import json
l=[x for x in open("$O/${T}_bench.log") if x.startswith("{")][-1]; d=json.loads(l); print(d["index_build"], d["value"])
PY
fi
