mkdir -p gpurun_out/r2i
python bench.py --full --init parafac-epc --no-e2e > gpurun_out/r2i/full_resnet18_epc.json 2> gpurun_out/r2i/full_resnet18_epc.err; tail -c 2500 gpurun_out/r2i/full_resnet18_epc.json; tail -3 gpurun_out/r2i/full_resnet18_epc.err
python bench.py --full --no-e2e > gpurun_out/r2i/full_resnet18_random.json 2> gpurun_out/r2i/full_resnet18_random.err; tail -c 1500 gpurun_out/r2i/full_resnet18_random.json; tail -3 gpurun_out/r2i/full_resnet18_random.err
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r2i/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2i/sanitizer_$tool.log
done
