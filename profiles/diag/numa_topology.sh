#!/bin/bash
# NUMA / PCIe topology of the box, for the multi-GPU end-to-end limiter
echo "== nodes"; for n in /sys/devices/system/node/node*; do echo "$(basename $n) cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo | awk '{print $4, $5}')"; done
echo "== allowed cpus / mems"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader | while IFS=, read i b; do b=$(echo $b | tr -d ' ' | tr 'A-Z' 'a-z'); echo "gpu $i $b numa_node=$(cat /sys/bus/pci/devices/${b:4}/numa_node 2>/dev/null)"; done
echo "== topo"; nvidia-smi topo -m 2>&1 | head -24
