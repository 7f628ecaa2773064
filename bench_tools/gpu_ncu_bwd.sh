#!/bin/bash
# launch list of the inversion step (32 images, D = 2 by default) + ncu --set full of the backward and the save-mode forward kernel
mkdir -p gpurun_out
export D=${D:-2} TARGETS=16 STEPS=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_inv_d$D.csv python bench_tools/bench_inversion.py > gpurun_out/ncu_inv_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused_backward_kernel|fused_forward_kernel' -s 8 -c 2 -f -o gpurun_out/prof_bwd_d$D python bench_tools/bench_inversion.py > gpurun_out/ncu_bwd.log 2>&1
tail -n 3 gpurun_out/ncu_bwd.log
python - <<'PY'
import csv, collections, os
D = os.environ.get("D", "2")
rows = [r for r in csv.reader(open(f"gpurun_out/launches_inv_d{D}.csv")) if len(r) > 10 and r[0].isdigit()]
# one fwd+bwd step = between consecutive fused_backward launches: aggregate the last 40 % of the file
agg = collections.OrderedDict()
for r in rows[len(rows) * 6 // 10:]:
    k = r[4][:60]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[-1]) / 1e3
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"{us:10.1f} us {n:4d}  {k}")
PY
