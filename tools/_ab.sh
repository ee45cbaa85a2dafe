cd $GRAFT_REPO_ROOT
bash tools/ncu_r2b.sh 2>&1 | tail -14
