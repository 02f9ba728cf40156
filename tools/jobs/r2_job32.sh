#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/debug/branches4.py 2>&1 | tail -20
