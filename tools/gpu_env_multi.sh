# several values of one environment switch on one config.  usage: bash tools/gpu_env_multi.sh TAG CFG VAR v1 v2 ...   ("-" = unset)
tag=$1; c=$2; var=$3; shift; shift; shift
for v in "$@"; do
  if [ "$v" = "-" ]; then e=""; else e="$var=$v"; fi
  env $e timeout 200 python bench.py --config $c --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_m${c}_$v.json 2> gpurun_out/${tag}_m${c}_$v.err || echo "$v failed"
done
python tools/stages.py gpurun_out/${tag}_m${c}_*.json
