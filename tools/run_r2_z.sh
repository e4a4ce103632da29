mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem.py -m gpu -x -q > gpurun_out/r2_z_stem.log 2>&1; echo "stem pytest (form 1) rc=$?"; tail -8 gpurun_out/r2_z_stem.log
RNB_NO_GRAPH=1 timeout 300 python tools/stem_epi_ab.py 256 0,1 > gpurun_out/stem_z.txt 2>&1; tail -4 gpurun_out/stem_z.txt
RNB_STEM_FUSED=0 RNB_NO_GRAPH=1 timeout 300 python tools/stem_epi_ab.py 256 0,1 > gpurun_out/stem_z_unfused.txt 2>&1; tail -4 gpurun_out/stem_z_unfused.txt
