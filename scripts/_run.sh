mkdir -p gpurun_out
P="python -m disentangledcolorization_b200.tools.conv_probe"
DISCO_TC_DEBUG=1 timeout 60 $P --cin 64 --cout 64 --hw 256 --batch 64 --iters 3 > gpurun_out/tl_64e.log 2>&1
tail -1 gpurun_out/tl_64e.log
