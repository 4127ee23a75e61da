run() { # tag env...
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --no-static --no-cpu-baseline > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/n2_$tag.json'));print('$tag',d['value'],d['ms_per_step'])"
}
run base X=1
run ch32 NCCL_MIN_NCHANNELS=32
run ch64 NCCL_MIN_NCHANNELS=64 NCCL_MAX_NCHANNELS=64
run simple32 NCCL_MIN_NCHANNELS=32 NCCL_PROTO=Simple
