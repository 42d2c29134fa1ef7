#!/bin/bash
# round-2 final evidence run (one B200): tests, smoke, both bench arms, ncu launch lists (C2 bench, C3 search) and
# full captures of the render-stage kernels and of the mesh kernels; the histogram / cluster kernels are unchanged
# since tools/gpu_final_r2.sh captured them (profiles/r02_hist_*, r02_cluster_*)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, "roofline frac", round(d["roofline"]["frac"], 3))
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
python tools/profile_run.py 0 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel|argmax|image_mode" -c 12 -f -o gpurun_out/prof_render python tools/profile_run.py 0 > gpurun_out/prof_ncu_render.log 2>&1
tail -1 gpurun_out/prof_ncu_render.log
python tools/exp_c3_profile.py 6 > gpurun_out/c3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 80 --csv --log-file gpurun_out/c3_launches.csv python tools/exp_c3_profile.py 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"mesh_vertices|mesh_raster|mesh_shade" -c 3 -f -o gpurun_out/prof_c3 python tools/exp_c3_profile.py 4 > gpurun_out/prof_ncu_c3.log 2>&1
tail -2 gpurun_out/c3_plain.log | head -1; tail -1 gpurun_out/prof_ncu_c3.log
./orbslam2_nmi_b200/_lib/ubench_atoms > gpurun_out/ubench_atoms.txt 2>&1; tail -2 gpurun_out/ubench_atoms.txt
python tools/exp_overflow_retry.py > gpurun_out/overflow_retry.txt 2>&1; tail -3 gpurun_out/overflow_retry.txt
