N=$1
for w in "wd5m-1vsall --batch 1024 --steps 5 --warmup 2" "wd5m-eval-complex --batch 512 --steps 5 --warmup 2" "wd5m-eval-transe --batch 128 --steps 3 --warmup 2"; do
  tools/run_bounded.sh 240 gpurun_out/scale_${N}.log python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench_extra.py --workload $w
  grep '^{"workload' gpurun_out/scale_${N}.log | cut -c1-330
done
