#!/bin/bash
# The gpurun calls behind the round-2 evidence under profiles/ (one call at a time; each block is one call's command line).
# Usage: tools/evidence_commands_r2.sh <block>   -- prints the command; run it with  /usr/local/graft/bin/gpurun --timeout N -- '<command>'
case "$1" in
  tests)      # GPUTEST: 152 passed, 1 skipped (2-GPU test; passes on a multi-GPU box)
    echo 'python -m pytest tests -m gpu -q; python -c "import __graft_entry__ as g; g.smoke()"' ;;
  bench)      # r02_bench_v15_{default,s4,stream16,reference}.json, r02_bench_v12_highres.json
    echo 'python bench.py --steps 10 --warmup 3; EPRECON_STREAMS=4 python bench.py --steps 6 --warmup 3; python bench.py --workload stream16 --steps 3 --warmup 3; python bench.py --workload highres --steps 5 --warmup 3; python bench.py --impl reference --steps 2 --warmup 1' ;;
  launches)   # r02_launches_v12_one_fragment.csv (+ tools/summarize_launches.py -> _summary.txt), r02_spconv_per_launch_v12.jsonl
    echo 'EPRECON_STREAMS=1 EPRECON_BENCH_SKIP_CPU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1; EPRECON_BENCH_DUMP=gpurun_out/spconv_per_launch.jsonl python bench.py --steps 10 --warmup 3' ;;
  ncu_conv)   # r02_spconv_hl_v5_level2_ncu_summary.txt + r02_spconv_hl_ncu_traffic.json (tools/ncu_summary.py --stalls --traffic-json ... --algorithmic 83919552)
    echo 'ncu --set full --clock-control none --import-source on -k regex:spconv_hl_cp -c 2 -f -o gpurun_out/r02_hl_prof python tools/probes/ncu_hl_one.py' ;;
  ncu_linear) # r02_linear_tile_ncu_summary.txt
    echo 'ncu --set full --clock-control none --import-source on -k regex:"linear_mma|spconv_kernel" -c 4 -f -o gpurun_out/r02_linear_prof python tools/probes/ncu_linear_one.py' ;;
  probes)     # r02_probe_hl_*_timing.json, r02_probe_linear_mma.json, r02_timeline_hl_*.json
    echo 'python tools/probes/probe_hl.py --part timing0; python tools/probes/probe_linear.py; python tools/probes/timeline_hl.py' ;;
  host)       # r02_host_profile_v13.txt, r02_bench_v13_sync_*.json, r02_bench_v13_cores*.json
    echo 'python tools/host_profile.py; for m in auto yield blocking; do EPRECON_BENCH_SKIP_CPU=1 EPRECON_SYNC=$m python bench.py --steps 10 --warmup 3; done; EPRECON_BENCH_SKIP_CPU=1 EPRECON_SYNC=yield taskset -c 0-1 python bench.py --steps 6 --warmup 3' ;;
  n8)         # gpurun --gpus 8: r02_bench_n8_v15.json, r02_bench_n4_v15.json (+ the v12 / v13 / v14 variants with EPRECON_SYNC / EPRECON_STREAMS / EPRECON_BENCH_NO_EXCHANGE)
    echo 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3; python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 8 --warmup 3; python -m pytest tests/test_dist_gpu.py -q' ;;
  *) echo "blocks: tests bench launches ncu_conv ncu_linear probes host n8" ;;
esac
