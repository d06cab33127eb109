run() { # envs...
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), 'ms/step')"
}
run A=1
run PS3D_NO_SCATTER_FENCE=1
run PS3D_P2P_CTAS=2
run PS3D_NO_FUSED_UPDATE=1
