#!/bin/bash
# ncu --set full captures of every kernel on the path (one GPU), run on the GPU box:  bash tools/r2_profile_all.sh [names...]
# Each report is reduced ON THE BOX to its raw + source CSV pages (gzip) and a text summary; the .ncu-rep files are deleted
# (gpurun_out/ travels back only below 64 MiB).
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launches to skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    $N -k regex:$rx --launch-skip $skip --launch-count 1 -o gpurun_out/r2_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
    ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2_$name.raw.csv.gz
    ncu -i gpurun_out/r2_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_$name.source.csv.gz
    python tools/summarize_ncu.py gpurun_out/r2_$name.ncu-rep "$@" > gpurun_out/r2_$name.summary.txt 2>&1
    if [ $name = tau_leap ] || [ $name = queue ]; then python tools/summarize_profile.py gpurun_out/r2_$name.ncu-rep 262144 gpurun_out/r2_${name}_ncu_full.txt "$LABEL" 48 > /dev/null 2>&1; fi
    python tools/top_lines.py gpurun_out/r2_$name.ncu-rep 40 > gpurun_out/r2_$name.lines.txt 2>&1
    rm -f gpurun_out/r2_$name.ncu-rep
    tail -n 1 gpurun_out/r2_ncu_$name.log
}
want=${@:-tau_leap sirsi pfr25 pfr100 sir_ode ssa loglik}
for w in $want; do
  case $w in
    tau_leap) cap tau_leap tau_leap 2 python tools/profile_kernel.py 262144 ;;
    queue)    cap queue tau_leap 2 python tools/profile_kernel.py 262144 259 ;;
    sirsi)    cap sirsi tau_leap 1 python tools/profile_configs.py sirsi ;;
    pfr25)    cap pfr25 ode_pfr 1 python tools/profile_configs.py pfr25 ;;
    pfr100)   cap pfr100 ode_pfr 1 python tools/profile_configs.py pfr100 1024 ;;
    sir_ode)  cap sir_ode ode_thread 1 python tools/profile_configs.py sir_ode ;;
    sir_radau) cap sir_radau ode_thread 1 python tools/profile_configs.py sir_radau ;;
    sir_bdf)  cap sir_bdf ode_thread 1 python tools/profile_configs.py sir_bdf ;;
    ssa)      cap ssa ssa_kernel 1 python tools/profile_configs.py ssa ;;
    loglik)   cap loglik loglik 1 python tools/profile_configs.py loglik ;;
  esac
done
ls -la gpurun_out/ | head -40
