"""TEST INFRASTRUCTURE: build a host-emulation library from libsis3d CUDA sources (kernels without warp intrinsics, TMA or
tcgen05), so that their index arithmetic, shared-memory staging and barriers run -- and can be race-checked with
ThreadSanitizer -- on a machine without a GPU.

    python tools/cuda_host_emu.py OUT.so SRC.cu [SRC.cu ...] [--tsan]

The source is rewritten textually: `kernel<<<grid, block, smem, stream>>>(args)` -> `emu_launch(kernel, dim3(grid), dim3(block),
args)` and `extern __shared__ T name[];` -> a static array; everything else is handled by the headers in csrc/emu_shims/
(host_emu.h: one CUDA block = blockDim std::threads meeting at a std::barrier for __syncthreads)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3d-sis_b200", "csrc")
LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(text):
    def launch(m):
        cfg = split_top(m.group(2))
        return f"emu_launch({m.group(1)}, dim3({cfg[0]}), dim3({cfg[1]}), "
    text = LAUNCH.sub(launch, text)
    text = re.sub(r"extern\s+__shared__\s+((?:__align__\(\d+\)\s+)?[\w ]+?)\s+(\w+)\s*\[\s*\]\s*;", r"static \1 \2[1 << 16];", text)
    return text


def build(out, sources, tsan=False):
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for src in sources:
            dst = os.path.join(tmp, os.path.basename(src).replace(".cu", "_emu.cpp"))
            with open(src) as f, open(dst, "w") as g:
                g.write(rewrite(f.read()))
            files.append(dst)
        glue = os.path.join(tmp, "emu_glue.cpp")
        with open(glue, "w") as g:
            g.write('#include "common.cuh"\nnamespace sis3d { unsigned long long g_launch_count = 0; }\n'
                    'extern "C" const char *sis3d_strerror(int code) { return code == 0 ? "ok" : "emulated libsis3d error"; }\n')
        cmd = ["g++", "-O1" if tsan else "-O2", "-g", "-std=c++20", "-DSIS3D_HOST_EMU", "-fPIC", "-shared", "-pthread",
               "-Wno-unknown-pragmas", "-Wno-attributes", "-ffp-contract=off", f"-I{os.path.join(CSRC, 'emu_shims')}", f"-I{CSRC}",
               f"-I{os.path.join(ROOT, 'include')}"] + (["-fsanitize=thread"] if tsan else []) + ["-o", out, glue] + files
        subprocess.check_call(cmd)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--tsan"]
    build(args[0], args[1:], tsan="--tsan" in sys.argv)
