for k in 3 7 11; do
MB_TC_PAIR_TRACE=gpurun_out/trace16_k${k}.txt MB_TC_PAIR_TRACE_F32=0 MB_TC_PAIR_TRACE_K=$k MB_TC_PAIR_TRACE_SKIP=7 timeout 200 python tools/profile_layers.py --precision f16tc --reps 2 2>&1 | tail -1
done
