#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 120 python tests/kernel_bench.py --batch 64 --chans 512 --hw 7 --only pw_fwd 2>&1 | tail -1
timeout 120 python tests/kernel_bench.py --batch 64 --chans 256 --hw 14 --only pw_fwd 2>&1 | tail -1
timeout 120 python tests/kernel_bench.py --batch 64 --only pw_fwd 2>&1 | tail -1
