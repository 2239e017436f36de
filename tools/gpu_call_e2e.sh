D=gpurun_out/${1:-e2e}; mkdir -p $D
( timeout 300 python -m pytest tests -m "e2e_gpu" -q -s > $D/pytest_e2e.log 2>&1; echo "pytest exit $?" >> $D/pytest_e2e.log )
tail -30 $D/pytest_e2e.log
