#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU, one kernel instance each), raw pages exported as CSV
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, tool arg, launch-skip
  timeout 300 $NCU -k "regex:$2" -s "$4" -c 1 -o gpurun_out/r02_ncu_$1 python tools/one_kernel.py "$3" > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_ncu_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/r02_ncu_$1.ncu-rep
  echo "$1: $(wc -l < gpurun_out/r02_ncu_$1_raw.csv) lines"
}
cap attn_fwd "attn_fwd_kernel" attn 1
cap attn_dkdv128 "attn_bwd_dkdv128" attn 1
cap wgrad_full "wgrad_full_kernel" wgrad_full 1
cap conv3x3 "gemm_bf16|conv" conv 1
cap optim "adamw_bf16_multi" optim 1
cap optim_big "adamw_bf16_multi" optim_big 1
cap lokr_rebuild "lokr_rebuild" lokr 1
cap lokr_factor "lokr_factor_grad" lokr 1
cap gelu "gelu_tanh_kernel" gelu 1
# launch list of the VAE encode bench (cheap: 445 launches per step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_vae.csv python bench.py --config vae_encode --steps 1 --warmup 3 > gpurun_out/r02_launches_vae.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_vae.csv 12 > gpurun_out/r02_launches_vae_summary.txt 2>&1
tail -14 gpurun_out/r02_launches_vae_summary.txt
# launch list of exactly one device-resident Flux step (profiler range inside bench.py)
STB_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_flux_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_launches_flux_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_flux_step.csv 40 > gpurun_out/r02_launches_flux_step_summary.txt 2>&1
head -30 gpurun_out/r02_launches_flux_step_summary.txt
