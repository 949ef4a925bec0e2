R=r02
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
cap ${R}_ncu_lif_bwd lif_bwd 1 1 python tools/ncu_targets.py lif_bwd
cap ${R}_ncu_conv_fwd gemm_kernel 2 1 python tools/ncu_gemm_targets.py conv_fwd
cap ${R}_ncu_conv_dgrad gemm_kernel 2 1 python tools/ncu_gemm_targets.py conv_dgrad
cap ${R}_ncu_lin_fwd gemm_kernel 2 1 python tools/ncu_gemm_targets.py lin_fwd
ls -la gpurun_out/${R}_ncu_lif_bwd.raw.csv gpurun_out/${R}_ncu_conv_fwd.raw.csv
