# A/B of an environment switch on the default build.  usage: bash tools/gpu_env_ab.sh TAG "cfgs" ENVVAR=VALUE   (runs each config without and with it)
tag=$1; cfgs=$2; sw=$3
mkdir -p gpurun_out
for c in $cfgs; do
  timeout 200 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_m${c}_base.json 2> gpurun_out/${tag}_m${c}_base.err || echo "base cfg $c failed"
  env $sw timeout 200 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_m${c}_switch.json 2> gpurun_out/${tag}_m${c}_switch.err || echo "switch cfg $c failed"
  python tools/stages.py gpurun_out/${tag}_m${c}_base.json gpurun_out/${tag}_m${c}_switch.json
done
