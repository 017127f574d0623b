#!/bin/bash
# Evidence set of one kernel generation, taken on a GPU box:  gpurun --timeout 1500 -- 'bash profiles/capture.sh r2'
#   bench lines (cfg2 default, cfg3, cfg4, reference arm), ncu --set full summaries for cfg2 / cfg3 / cfg4, the launch list of the
#   default bench command, per-phase clock timelines.  Everything lands in gpurun_out/<tag>_*; copy what is kept into profiles/.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --workload cfg3 --no-binary > $out/${tag}_bench_n1_cfg3.json 2> $out/${tag}_bench_n1_cfg3.err
python bench.py --workload cfg4 --no-binary > $out/${tag}_bench_n1_cfg4.json 2> $out/${tag}_bench_n1_cfg4.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
for wl in cfg2 cfg3 cfg4; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:hfg_estep_v3 -s 4 -c 1 -f -o $out/${tag}_${wl}_estep \
        python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-binary > $out/${tag}_${wl}_ncu.log 2>&1
    {
        echo "ncu --set full --clock-control none --import-source on -k regex:hfg_estep_v3 -s 4 -c 1   python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-binary"
        echo "(ncu serialises and replays the launch: compare shares, not absolute time; the CUDA-event time per iteration is in the bench line of the same workload)"
        echo
        python profiles/ncu_summary.py $out/${tag}_${wl}_estep.ncu-rep
    } > $out/${tag}_ncu_${wl}_summary.txt 2>&1
    [ $wl != cfg2 ] && rm -f $out/${tag}_${wl}_estep.ncu-rep
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-binary > $out/${tag}_launches.log 2>&1
for wl in cfg2 cfg3 cfg4; do python tools/quad_phases.py $wl; done > $out/${tag}_phases.txt 2>&1
{
    echo "== scout warps off (HFG_DBG=32), config 2"; HFG_DBG=32 python tools/quad_phases.py cfg2 | tail -1
    echo "== tree rounds only (HFG_DBG=16), config 2"; HFG_DBG=16 python tools/quad_phases.py cfg2 | tail -1
    echo "== rate fit: predicted against tree rounds, same bits"; python tools/fit_check.py small cfg2 cfg4 | cut -c1-600
} > $out/${tag}_mstep_tail.txt 2>&1
python tools/batch_bench.py 7.5e7 3e8 1.5e9 3e9 > $out/${tag}_batch_bench.txt 2>&1
{
    for b in lat_bench pred_bench mio_bench a_bench launch_bench; do
        [ -x tools/$b ] && { echo "== tools/$b"; timeout 120 ./tools/$b; }
    done
} > $out/${tag}_microbench.txt 2>&1
echo capture done
