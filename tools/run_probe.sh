# Next experiment (prepared, not yet run): does releasing a ring stage one k-tile late cure the TMA kernel's failure when it
# shares SMs with other kernels?  (baseline GEGP_TMA_SHARED_SM=1: 18 of 30 repetitions deviate, solo config: 0.)
for v in 1 2 0; do
  GEGP_TMA_SHARED_SM=$v GEGP_NO_LOOKAHEAD=1 REPRO_LOAD=20 python tools/repro_probe.py 1000 20 0 30 2>&1 | tail -1
done
# if 2 is clean: speed of the shared configuration with the late release against the solo one
GEGP_TMA_SHARED_SM=2 python tools/quick_probe.py 500,10 1000,20 2>&1 | tail -2
python tools/quick_probe.py 500,10 1000,20 2>&1 | tail -2
