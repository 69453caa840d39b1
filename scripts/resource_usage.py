"""Per-kernel resource usage of the shipped library (`cuobjdump --dump-resource-usage`): registers, stack (spill
frames), static shared memory.  Writes profiles/r02_resource_usage.json; run after a build, no GPU needed."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "quantool_b200", "lib", "libquantool_b200.so")


def main():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function (\S+):", txt)), capture_output=True,
                           text=True).stdout.splitlines()
    rows = []
    for name, m in zip(names, re.finditer(r"Function \S+:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt)):
        rows.append({"kernel": name, "registers": int(m.group(1)), "stack_bytes": int(m.group(2)),
                     "static_shared_bytes": int(m.group(3)), "local_bytes": int(m.group(4))})
    rows.sort(key=lambda r: (-r["stack_bytes"], -r["registers"]))
    out = {"library": os.path.relpath(SO, ROOT), "kernels": len(rows),
           "with_stack_frame": [r for r in rows if r["stack_bytes"] or r["local_bytes"]],
           "max_registers": rows and max(r["registers"] for r in rows), "all": rows}
    with open(os.path.join(ROOT, "profiles", "r02_resource_usage.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(rows), "kernels;", len(out["with_stack_frame"]), "with a stack frame")
    for r in out["with_stack_frame"]:
        print("  ", r["stack_bytes"], r["registers"], r["kernel"][:140])


if __name__ == "__main__":
    main()
