# role timelines of CTA 0 of one fp32-input pair launch (k = 3 and k = 11), new (RT) and round-2 kernel; see tools/trace_pair_report.py
for rt in 1 0; do for k in 3 11; do
MB_TC_PAIR_RT=$rt MB_TC_PAIR_TRACE=gpurun_out/trace_rt${rt}_k${k}.txt MB_TC_PAIR_TRACE_K=$k MB_TC_PAIR_TRACE_SKIP=7 timeout 200 python tools/profile_layers.py --precision f16tc --reps 2 2>&1 | tail -1
done; done
