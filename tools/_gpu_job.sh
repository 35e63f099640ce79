cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
timeout 90 python tools/tc_timeline.py > gpurun_out/r2t_timeline.txt 2>&1; echo "timeline rc=$?"; grep "step end\|forward step" gpurun_out/r2t_timeline.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
bash tools/run_configs.sh > gpurun_out/r2t_configs.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_configs.jsonl'):
    d=json.loads(l); print(d['config']['global_batch'], d['config']['workload'][:40], '| value %.1fM ms %.3f | e2e %.1fM upd/s %.1f | frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['updates_per_s'], d['roofline']['frac']))
PY
