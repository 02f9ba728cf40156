#!/bin/bash
cd "$GRAFT_REPO_ROOT"
cd "$GRAFT_REPO_ROOT" && timeout 150 python -m pyrddlgym_jax_b200 tune Reservoir_Continuous 1 slp -t 1 -i 3 -w 1 --train-seconds 1 -f /tmp/best.cfg 2>&1 | tail -12; ls -la /tmp/best.cfg /tmp/gp_slp_* 2>&1 | tail -3
