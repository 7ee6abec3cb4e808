RL_EXP_PLANAR=1 RL_FWD_DBG=128 timeout 120 python tools/per_launch.py dense_fwd 2>&1 | tail -1
