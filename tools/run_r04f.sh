#!/bin/bash
cd $GRAFT_REPO_ROOT
echo "folded equivalent of 64 x 512 x 7x7 (one image of 58x58 = 3364 pixels):"
timeout 120 python tests/kernel_bench.py --batch 1 --chans 512 --hw 58 --only pw_fwd,pw_dgrad,pw_wgrad 2>&1 | tail -3
echo "current 64 x 512 x 7x7:"
timeout 120 python tests/kernel_bench.py --batch 64 --chans 512 --hw 7 --only pw_fwd,pw_dgrad,pw_wgrad 2>&1 | tail -3
echo "folded equivalent of 64 x 256 x 14x14 (one image of 112x112 = 12544 pixels):"
timeout 120 python tests/kernel_bench.py --batch 1 --chans 256 --hw 112 --only pw_fwd,pw_dgrad,pw_wgrad 2>&1 | tail -3
echo "current 64 x 256 x 14x14:"
timeout 120 python tests/kernel_bench.py --batch 64 --chans 256 --hw 14 --only pw_fwd,pw_dgrad,pw_wgrad 2>&1 | tail -3
