python scripts/pred50_check.py > gpurun_out/pred50.log 2>&1
python scripts/phase_profile.py adam_cfg4 20 > gpurun_out/phase_adam.log 2>&1
python scripts/phase_profile.py sgld_cfg1x 20 > gpurun_out/phase_cfg1x.log 2>&1
tail -20 gpurun_out/pred50.log gpurun_out/phase_adam.log gpurun_out/phase_cfg1x.log
