nproc; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" ; nvidia-smi topo -m | head -14
run() { tag=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --e2e-only --steps 200 --warmup 10 > gpurun_out/a5_e2e8_$tag.json 2> gpurun_out/a5_e2e8_$tag.err; tail -c 400 gpurun_out/a5_e2e8_$tag.json; echo; }
run pin SMOS_X=1
run nopin SMOS_NO_PIN=1
run noh2d SMOS_E2E_NO_H2D=1
