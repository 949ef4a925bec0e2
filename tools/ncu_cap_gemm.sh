cap() {  # name, kernel regex, skip, target
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/ncu_gemm_targets.py $4 > /dev/null 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap r02_ncu_conv_wgrad wgrad_kernel 2 conv_wgrad
cap r02_ncu_lin_wgrad wgrad_kernel 2 lin_wgrad
cap r02_ncu_lin_fwd gemm_kernel 2 lin_fwd
cap r02_ncu_conv_fwd gemm_kernel 2 conv_fwd
ls -la gpurun_out/ | grep r02_ncu
