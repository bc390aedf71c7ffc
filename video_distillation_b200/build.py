"""In-tree build of libvd_b200.so (sm_100a only) with plain nvcc.

    python -m video_distillation_b200.build [--force] [--verbose]

The shared object is written next to this file so that it travels with the repo snapshot to
the GPU box; it has no dependency on torch or libcuda (runtime API only, statically linked).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvd_b200.so')
SOURCES = ['api.cu', 'simt_conv.cu', 'pointwise.cu', 'compose_tiled.cu', 'compose_tma.cu', 'tc_conv.cu', 'tc_pack.cu', 'tc_bwd.cu', 'tc_trio.cu']
PROBE_SOURCES = SOURCES + ['tc_probe.cu']      # + -DVD_PROBE: scripts/libvd_b200_probe.so (tuning tools only, see scripts/_probe_lib.py)
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC']


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'vd_b200.h'))
    if not force and os.path.exists(LIB) and not any(_newer(d, LIB) for d in deps):
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, 'build', os.path.basename(s) + '.o')
        objs.append(o)
        if force or any(_newer(d, o) for d in [s] + [d for d in deps if not d.endswith('.cu')]):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {os.path.basename(s)} (exit {p.returncode})\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed building libvd_b200.so')
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return LIB


def build_probe(out_path, verbose=False):
    """The same sources + csrc/tc_probe.cu compiled with -DVD_PROBE into `out_path`: the tuning / bring-up entry points
    (vd_tc_probe, vd_tc_mma_rate*, vd_tc_set_profile_buffer) live ONLY there."""
    srcs = [os.path.join(CSRC, s) for s in PROBE_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    if os.path.exists(out_path) and not any(_newer(d, out_path) for d in deps):
        return out_path
    bdir = os.path.join(HERE, 'build', 'probe')
    os.makedirs(bdir, exist_ok=True)
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + '.o')
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC] + FLAGS + ['-DVD_PROBE', '-c', s, '-o', o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {s}:\n{out}')
        if verbose:
            sys.stderr.write(out)
    r = subprocess.run([NVCC, '-shared', '-o', out_path] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return out_path


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
