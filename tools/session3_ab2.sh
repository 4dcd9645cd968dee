# A/B of march CTA shapes (session 3): 3 CTAs x 2 warpgroups (default) vs 2 x 3 vs 1 x 6 - same 24 warps per SM
for rep in 1 2; do
for lib in libnmr.so libnmr_g3c2.so libnmr_g6c1.so; do
  bash tools/ab1.sh "rep$rep $lib" nerf-glasses_b200/$lib
  bash tools/ab1.sh "rep$rep $lib zoom4" nerf-glasses_b200/$lib --zoom 4
  bash tools/ab1.sh "rep$rep $lib zoom4 translucent" nerf-glasses_b200/$lib --zoom 4 --regime translucent --steps 20
done
done > gpurun_out/s3_ab2.txt 2>&1
cat gpurun_out/s3_ab2.txt
