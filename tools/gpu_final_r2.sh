#!/bin/bash
# round-2 evidence run: tests, smoke, both bench arms, ncu launch list + full captures, ubench, stress frames
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, "roofline frac", round(d["roofline"]["frac"], 3))
PY
# launch list of the same command line (only after it exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
# full captures of one steady-state search (profile_run.py brackets its last search with cudaProfilerStart/Stop)
python tools/profile_run.py 0 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel|argmax|image_mode" -c 12 -f -o gpurun_out/prof_render python tools/profile_run.py 0 > gpurun_out/prof_ncu_render.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"joint_hist" -c 1 -f -o gpurun_out/prof_hist python tools/profile_run.py 0 > gpurun_out/prof_ncu_hist.log 2>&1
tail -1 gpurun_out/prof_ncu_render.log; tail -1 gpurun_out/prof_ncu_hist.log
./orbslam2_nmi_b200/_lib/ubench_atoms > gpurun_out/ubench_atoms.txt 2>&1; tail -2 gpurun_out/ubench_atoms.txt
for f in uniform sky constant; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --frame $f > gpurun_out/bench_frame_$f.json 2> gpurun_out/bench_frame_$f.err
  python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_frame_{sys.argv[1]}.json"))
print("frame", sys.argv[1], round(d["value"]), "evals/s", {k: round(x, 3) for k, x in d["stage_ms"].items()})
PY
done
# single evaluation (config 1): launch list and a full capture of the cluster kernel
python tools/exp_c1_launches.py > gpurun_out/c1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c1_launches.csv python tools/exp_c1_launches.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"joint_hist_score_cluster" -s 2 -c 1 -f -o gpurun_out/prof_cluster python tools/exp_c1_launches.py > gpurun_out/prof_ncu_cluster.log 2>&1
tail -1 gpurun_out/prof_ncu_cluster.log
