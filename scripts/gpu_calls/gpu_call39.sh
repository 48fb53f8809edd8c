#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py pair_c128_128 pair_c64_256_stats pair_s2 pair_c320 perfp_s2_64_128_128_pair perfp_s2_64_128_128_solo perfp_s2_32_64_128_pair perfs2_32_64_128_b8 > gpurun_out/bringup_pair2.log 2>&1; echo "bringup_exit=$?"; grep -E "PASS|FAIL|failures" gpurun_out/bringup_pair2.log | sed 's/box .*| //'
