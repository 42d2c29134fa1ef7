#!/bin/bash
# copy the outputs of tools/gpu_final_r2b.sh from gpurun_out/ into profiles/ (summaries only)
set -e
cp gpurun_out/bench_default.json profiles/r02_bench_default.json
cp gpurun_out/bench_reference.json profiles/r02_bench_reference.json
cp gpurun_out/launches.csv profiles/r02_launches.csv
python tools/summarize_launches.py gpurun_out/launches.csv profiles/r02_launches_summary.txt "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs" > /dev/null
cp gpurun_out/c3_launches.csv profiles/r02_c3_launches.csv
python tools/summarize_launches.py gpurun_out/c3_launches.csv profiles/r02_c3_launches_summary.txt "python tools/exp_c3_profile.py 4 (one steady-state C3 search)" > /dev/null
python tools/summarize_ncu.py gpurun_out/prof_render.ncu-rep profiles/r02_render > /dev/null
python tools/summarize_ncu.py gpurun_out/prof_c3.ncu-rep profiles/r02_c3 > /dev/null
cp gpurun_out/ubench_atoms.txt profiles/r02_ubench_shared_atomics.txt
cp gpurun_out/overflow_retry.txt profiles/r02_overflow_retry.txt
tail -4 gpurun_out/pytest_gpu.log > profiles/r02_pytest_gpu.txt
tail -1 gpurun_out/smoke.log > profiles/r02_smoke.txt
python tools/sass_excerpts.py > /dev/null
python - <<'PY'
import json
d = json.load(open("profiles/r02_bench_default.json"))
with open("profiles/r02_frames.txt", "w") as f:
    f.write("# C2 search on the other camera frames (configs.C2_frames of profiles/r02_bench_default.json; hist_path 2 = the build with hot-bin skipping)\n")
    for k, v in d["configs"]["C2_frames"].items():
        f.write(f"frame {k}: search {v['search_ms']:.3f} ms, histogram {v['hist_ms']:.3f} ms, {v['evals_per_s']:.0f} evals/s" + (f", hist_path {v['hist_path']}" if "hist_path" in v else "") + "\n")
print(open("profiles/r02_frames.txt").read())
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable_frame"]["value"])
for k, v in d["configs"].items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("search_ms", "evals_per_s", "ms_per_frame_mean", "ms_per_frame_median", "ms_per_frame_p99", "ms_per_frame_max", "search_ms_wall", "search_ms_device", "eval_pair_call_us")})
PY
