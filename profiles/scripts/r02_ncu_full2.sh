#!/bin/bash
# the SECOND invocation of the level kernels (the first node level's assign / resolve, the second partition
# pass, the first real node_insert) - the first invocations are in r02_ncu_full_3100mbp.md
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --kernel-id '::regex:^(assign_kernel|resolve_kernel|node_insert_kernel|.*partition_kernel|count_kernel).*:2' -f -o /tmp/r02_full_2nd \
    python profiles/scripts/r02_ncu_target.py build > gpurun_out/r02_full_2nd_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_full_2nd_ncu.log
python profiles/scripts/summarize_ncu.py /tmp/r02_full_2nd.ncu-rep > gpurun_out/r02_ncu_full_3100mbp_second.md
cat gpurun_out/r02_ncu_full_3100mbp_second.md | cut -c1-300
