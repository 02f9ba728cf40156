#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 120 python -m pyrddlgym_jax_b200 plan Reservoir_Continuous 1 slp -e 1 --train-seconds 2 2>&1 | tail -4
timeout 120 python -m pyrddlgym_jax_b200 plan Pomdp_Toy 1 drp -e 2 --train-seconds 2 2>&1 | tail -3
timeout 120 python -m pyrddlgym_jax_b200 plan Quadcopter 1 replan -e 1 --train-seconds 0.1 2>&1 | tail -3
