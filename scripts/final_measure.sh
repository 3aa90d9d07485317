#!/bin/bash
# Round-end measurement pass on one B200 (run through gpurun); everything lands in gpurun_out/final/.
# usage: bash scripts/final_measure.sh [skip_ncu]
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cut -c1-400 $O/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
python scripts/bench_warp.py > $O/warp_config3.json 2> $O/warp_config3.err; cat $O/warp_config3.json
python scripts/bench_clip.py > $O/clip_config5.json 2> $O/clip_config5.err; cut -c1-400 $O/clip_config5.json
[ -n "$1" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --headline-only > $O/ncu_bench.log 2>&1; tail -1 $O/ncu_bench.log | cut -c1-200
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv --csv --log-file $O/conv_traffic.csv python scripts/layer_paths.py 64 > $O/paths.txt 2> $O/paths.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:k_(visibility|plane_gate|solve|warp_rows)' --csv --log-file $O/warp_config3_launches.csv python scripts/bench_warp.py 16384 1 > /dev/null 2>&1
ls -la $O
