#!/usr/bin/env bash
# How every number under profiles/ was produced (run from the repo root on a B200 box, e.g. through gpurun).
# Nothing printed under ncu is used as a bench value: the bench lines come from the plain runs.
set -euo pipefail
mkdir -p gpurun_out

# --- 1 GPU: default line (value, serial_value, roofline, e2e, cpu_baseline) and the reference arm
python bench.py                                   > gpurun_out/bench_default.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log

# --- variants: bf16 output + bf16 embeddings (tcgen05 heads), the other BASELINE shapes, letterbox, train pipeline
python bench.py --out-dtype bf16 --emb-dtype bf16 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e
for w in cfg3_1080p_20 cfg2_multitask_256 cfg4_vit_5heads cfg1_single_224; do
  python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-e2e
done
python bench.py --resize letterbox --steps 50 --warmup 5 --no-cpu-baseline --no-e2e
python bench.py --train-aug --steps 50 --warmup 5 --no-cpu-baseline --no-e2e
python bench.py --train-aug --resize letterbox --steps 50 --warmup 5 --no-cpu-baseline --no-e2e

# --- ncu: launch list (share of the step) and one full capture of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
ncu --set full --clock-control none --import-source on -k regex:"k1_crop|k2_heads_forward|k2_heads_dw" -c 3 \
    -o gpurun_out/k1_k2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
#   read back with:  ncu -i gpurun_out/k1_k2.ncu-rep --page details --csv   (and --page raw / --page source)

# --- N GPUs (one rank per GPU; K4' peer transport by default, --allreduce nccl for the NCCL baseline)
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/bench_${n}gpu.log || true
done

# --- parity
python -m pytest tests -x -q -m gpu
python profiles/tools/hsv_gpu_probe.py

# --- round 2 (profiles/r2/)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-api --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv $B
ncu --set full --clock-control none --import-source on -k regex:"k1_crop|k2_fused" -s 12 -c 2 -o gpurun_out/k1_k2fused -f $B
ncu --set full --clock-control none --import-source on -k regex:"k1_crop" -s 6 -c 1 -o gpurun_out/k1_aug -f $B --train-aug
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:k1_crop -s 4 -c 2 --csv --log-file gpurun_out/k1_crop_major_dram.csv $B
python profiles/tools/k2_bench.py; NKBK_FUSED_TIMING=1 python profiles/tools/k2_phases.py      # the fused heads step alone
python profiles/tools/engine_profile.py; python profiles/tools/engine_trace.py                # the API leg on the host / device
python profiles/tools/h2d_ceiling.py                                                          # (under torchrun for N > 1)
python bench.py --workload cfg5_shard8 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-variants --no-api
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712"
$T bench.py --gpus 8 --steps 50 --warmup 5 --no-e2e                                           # weak + strong + parity_check
$T profiles/tools/k4_phases.py                                                                # exchange phases, NCCL beside it
NKBK_K1_TIMING=1 python profiles/tools/k1_timeline.py --workload cfg5_1080p_64x64             # per-CTA timeline of one K1 launch
NKBK_K1_TIMING=1 python profiles/tools/k1_timeline.py --workload cfg5_shard8
python profiles/tools/epoch_profile.py                                                        # host time of a warm API epoch outside the loop
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712"
$T4 bench.py --gpus 4 --steps 50 --warmup 5; $T4 profiles/tools/h2d_ceiling.py                # 4-GPU line + H2D ceiling
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29679 tests/engine_dist_check.py
