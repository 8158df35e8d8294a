#!/bin/bash
# Multi-GPU measurement suite (one box, N GPUs): PCIe ceilings with all ranks copying, then the three bench
# workloads, every one launched exactly as the driver launches bench.py.   usage: tools/multi_gpu_suite.sh N [what...]
N=$1; shift
WHAT=${@:-probe corpus train segments}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "$N" = "1" ] && RUN="python"
mkdir -p gpurun_out
P=29520
for w in $WHAT; do
  P=$((P+1))
  [ "$N" = "1" ] && PORT="" || PORT="--master-port $P"
  case $w in
    probe)    timeout 300 $RUN $PORT tools/pcie_probe_multi.py gpurun_out/r02_pcie_probe_$N.json > gpurun_out/probe_$N.log 2>&1; tail -4 gpurun_out/probe_$N.log | cut -c1-400;;
    corpus)   timeout 400 $RUN $PORT bench.py --gpus $N --workload corpus --steps 5 --warmup 3 > gpurun_out/r02_corpus_$N.json 2> gpurun_out/corpus_$N.err; tail -c 300 gpurun_out/corpus_$N.err; head -c 250 gpurun_out/r02_corpus_$N.json; echo;;
    train)    timeout 600 $RUN $PORT bench.py --gpus $N --workload train --steps 8 --warmup 3 > gpurun_out/r02_train_$N.json 2> gpurun_out/train_$N.err; tail -c 300 gpurun_out/train_$N.err; head -c 250 gpurun_out/r02_train_$N.json; echo;;
    segments) timeout 600 $RUN $PORT bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_segments_$N.json 2> gpurun_out/segments_$N.err; tail -c 300 gpurun_out/segments_$N.err; head -c 250 gpurun_out/r02_segments_$N.json; echo;;
  esac
done
