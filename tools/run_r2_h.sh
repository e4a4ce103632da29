mkdir -p gpurun_out
python tools/dual_ab.py resnet50 256 > gpurun_out/dual_r50.txt 2>&1; cat gpurun_out/dual_r50.txt
RNB_NO_PDL=1 python tools/dual_ab.py resnet50 256 > gpurun_out/dual_r50_nopdl.txt 2>&1; cat gpurun_out/dual_r50_nopdl.txt
python tools/dual_ab.py resnet152 128 > gpurun_out/dual_r152.txt 2>&1; cat gpurun_out/dual_r152.txt
RNB_NO_PDL=1 python tools/dual_ab.py resnet152 128 > gpurun_out/dual_r152_nopdl.txt 2>&1; cat gpurun_out/dual_r152_nopdl.txt
