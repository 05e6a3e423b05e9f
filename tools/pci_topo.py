"""Where each visible GPU hangs in the PCI tree, from sysfs only (no CUDA work): bus id, the resolved sysfs path (every
upstream bridge is a path component), the root complex, NUMA node, link speed / width.  Used to decide how ranks should be
spread over GPUs when fewer ranks than GPUs run (bench.py: GPUs behind different host bridges do not share an upstream link).

    python tools/pci_topo.py
"""
import json
import os
import subprocess


def read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def main():
    q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"], capture_output=True, text=True).stdout
    out = []
    for line in q.strip().splitlines():
        idx, bus, gen, width = [x.strip() for x in line.split(",")]
        bdf = bus.lower()[4:] if len(bus) > 12 else bus.lower()   # nvidia-smi prints an 8-digit domain
        p = f"/sys/bus/pci/devices/{bdf}"
        real = os.path.realpath(p)
        out.append({"gpu": int(idx), "bus_id": bus, "sysfs": real, "root": real.split("/")[3] if real.count("/") > 3 else None,
                    "depth": real.count("/") - 3, "numa_node": read(p + "/numa_node"), "link": f"gen{gen} x{width}",
                    "max_link_speed": read(p + "/max_link_speed"), "iommu_group": os.path.basename(os.path.realpath(p + "/iommu_group")) if os.path.exists(p + "/iommu_group") else None})
    print(json.dumps({"gpus": out, "lspci_tree": subprocess.run("lspci -tv 2>/dev/null | head -80", shell=True, capture_output=True, text=True).stdout}, indent=1))


if __name__ == "__main__":
    main()
