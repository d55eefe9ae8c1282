# compute-sanitizer over the GPU tests (memcheck, racecheck, initcheck).  usage: bash tools/gpu_sanitize.sh TAG   (under gpurun)
tag=${1:-san}
mkdir -p gpurun_out
SEL='icosphere_cfg1 or single_triangle or ground_plane or occlusion_wall or cull_stress or resize or device_draw_list or resolve or hiz_only or golden or known_answer or pipeline or malformed'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/${tag}_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/${tag}_$tool.log
  grep -E "SUMMARY|passed|failed|^exit" gpurun_out/${tag}_$tool.log | tail -3
done
