mkdir -p gpurun_out
python tools/sampler_ab.py resnet50 256 > gpurun_out/sampler_r50.txt 2>&1; cat gpurun_out/sampler_r50.txt
python tools/sampler_ab.py resnet152 128 > gpurun_out/sampler_r152.txt 2>&1; cat gpurun_out/sampler_r152.txt
