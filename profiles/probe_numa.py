import os, glob, torch
print("affinity", sorted(os.sched_getaffinity(0)))
print("nodes", [ (p, open(p+"/cpulist").read().strip()) for p in sorted(glob.glob("/sys/devices/system/node/node*"))])
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bid = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    d = "/sys/bus/pci/devices/" + bid
    def rd(f):
        try: return open(d + "/" + f).read().strip()
        except Exception as e: return repr(e)
    print(i, bid, "numa", rd("numa_node"), "cpus", rd("local_cpulist"), rd("current_link_speed"), rd("current_link_width"))
os.system("nvidia-smi topo -m 2>&1 | head -30; lscpu | head -30; free -g | head -3")
