for v in 1 0; do RL_XP=$v timeout 120 python tools/per_launch.py dz_prepare 2>&1 | tail -1; done
