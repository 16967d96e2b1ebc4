for v in base a b ab; do
  if [ $v = base ]; then L=feltor_b200/libdgb200.so; else L=feltor_b200/exp/libdgb200_$v.so; fi
  DGB200_LIB=$PWD/$L timeout 300 python tools/pcg_stage_times.py 512 1024 2>&1 | grep "auto" | sed "s/^/$v /" >> gpurun_out/exp_libs.txt
done
