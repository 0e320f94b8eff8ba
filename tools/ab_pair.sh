run() { name=$1; shift; env "$@" timeout 200 python tools/profile_layers.py --precision f16tc --out gpurun_out/ab_$name.tsv 2>&1 | tail -1; }
run A MB_TC_PAIR_RT=0 MB_TC_PAIR_CV=2
run B MB_TC_PAIR_RT=0 MB_TC_PAIR_CV=4
run C MB_TC_PAIR_RT=0 MB_TC_PAIR_CV=4 MB_TC_PAIR_EW16=0
run D MB_TC_PAIR_RT=1
run E MB_TC_PAIR_RT=2
