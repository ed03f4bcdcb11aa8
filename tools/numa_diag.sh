python - <<'PY'
import torch, os, glob
p = torch.cuda.get_device_properties(0)
print([a for a in dir(p) if 'pci' in a.lower() or 'uuid' in a.lower()])
for a in ('pci_bus_id','pci_device_id','pci_domain_id'):
    print(a, getattr(p, a, None))
print(os.sched_getaffinity(0))
PY
nvidia-smi --query-gpu=index,pci.bus_id --format=csv
nvidia-smi topo -m 2>&1 | head -20
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/vendor); fi; done | head -12
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head
