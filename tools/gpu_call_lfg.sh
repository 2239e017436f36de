# LFG decoder first validation: stage-by-stage diagnosis, then the staged LFG tests together with the regular GPU suite
D=gpurun_out/${1:-lfg}; mkdir -p $D
timeout 300 python tools/lfg_diag.py > $D/lfg_diag.log 2>&1; echo "exit $?" >> $D/lfg_diag.log
( timeout 500 python -m pytest tests -m "lfg_gpu" -q -s > $D/pytest_lfg.log 2>&1; echo "pytest exit $?" >> $D/pytest_lfg.log )
( timeout 500 python -m pytest tests -m "gpu" -q > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
cat $D/lfg_diag.log | tail -30; grep -E "passed|failed|exit|x tol|Error" $D/pytest_lfg.log | tail -15; tail -3 $D/pytest_gpu.log
