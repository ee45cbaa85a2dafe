cd $GRAFT_REPO_ROOT
python tools/parity_report.py > gpurun_out/parity_r2_v2.txt 2>&1
tail -8 gpurun_out/parity_r2_v2.txt | cut -c1-600
