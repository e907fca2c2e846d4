#!/bin/bash
# Box diagnostics (run under gpurun): cores, memory, NUMA, GPU topology.
mkdir -p gpurun_out
{
  echo "== nproc"; nproc
  echo "== free -g"; free -g
  echo "== cgroup mem"; cat /sys/fs/cgroup/memory.max 2>/dev/null
  echo "== lscpu"; lscpu | head -30
  echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null; for n in /sys/devices/system/node/node*; do echo $n; cat $n/cpulist; done
  echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ -f $d/vendor ] && grep -q 0x10de $d/vendor && [ "$(cat $d/class)" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
  echo "== nvidia-smi"; nvidia-smi
  echo "== topo"; nvidia-smi topo -m
  echo "== affinity"; taskset -p $$
  echo "== ulimit -l"; ulimit -l
} > gpurun_out/box_info.txt 2>&1
