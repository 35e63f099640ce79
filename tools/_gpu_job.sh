cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tc or golden" > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2r_tests.log
bash tools/run_configs.sh
python - <<'PY'
import json
for l in open('gpurun_out/r2_configs.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:70], '| value %.1fM ms %.3f | e2e %.1fM upd/s %.1f | frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['updates_per_s'], d['roofline']['frac']))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_rollout_kernel -s 6 -c 1 -o gpurun_out/r2_prof_rollout_main python tools/ncu_target.py 2 > gpurun_out/r2r_ncu_full.log 2>&1; echo "ncu full rc=$?"
