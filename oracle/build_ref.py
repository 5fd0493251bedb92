"""Builds oracle/_ref/: the REFERENCE'S OWN harness, engine and GCN operator code, compiled unchanged from where it lies under
/root/reference, against
  * cognn_b200/host/shim/include/   drop-ins for the five Task-Worker headers the reference includes but does not ship
                                    (SCIHarness.h, ObliviousMapper.h, SecureAggregation.h, TaskUtil.h, TaskqHandler.h), on the C ABI;
  * tests/ref_standins/             minimal stand-ins for Boost.Serialization / Iostreams and cryptoTools' Network (test only).

  oracle/_ref/gcn-optimize            algo_kernels/common_harness/harness.cpp + vertex_centric/optimize-gcn/{gcn,kernel_harness}.h
  oracle/_ref/gcn-inference-optimize  the same with vertex_centric/optimize-gcn-inference
  oracle/_ref/gcn-original            the same with vertex_centric/original-gcn (needs the fused primitives: built if they compile)

What this is for: (1) the drop-in proof -- the reference's main(), SSEdgeCentricAlgoKernel::operator() and GCNEdgeCentricAlgoKernel
hooks run as written on the B200 library (tests/test_gpu_reference_dropin.py); (2) pinning the epoch oracle: the loss / accuracy
lines this binary prints must equal oracle/epoch.py's (tests/test_reference_dropin.py runs it on the CPU mock of the C ABI).
No reference source is copied; outputs go only into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).

Flags.  The reference's engine has latent faults that decide how it can be built (none of them is in code this repo replaces):
  * ss_vertex_centric_algo_kernel.h:925 hands the helper (BOB) threads a reference to a barrier that is a LOCAL of
    runAlgoKernelServer, which returns while they run (ASan: stack-use-after-return; a plain -O2 build reuses the dead frame
    and glibc aborts in pthread_mutex_lock, the reference's own -O0 CMake build survives by stack-layout luck).  The build
    below (-O1 -fwhole-program with the caller-size limits lifted) makes GCC inline that function, called once, into
    SSEdgeCentricAlgoKernel::operator(), whose frame outlives the threads -- no source is touched.
  * graph.h:625 reads edges_[size()-2] when the first edge is inserted (ASan: heap-buffer-overflow; harmless).
  * ssk.h:734-763 lets the ALICE threads of non-primary peers read gs.localVertexSvv while the primary thread is still inside
    PreScatterComp, and ssk.h:1067 lets a helper thread that runs one iteration ahead swap gs.localUpdateSvvs[i] under the
    owner thread's GatherComp: races that the WAN latency of a real deployment hides and a loopback run does not.  Two-party
    runs complete in ~85 % of the attempts (the tests retry); with more parties the first epoch usually completes and matches the
    oracle, later epochs are unreliable."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("COGNN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
VARIANTS = {"gcn-optimize": "optimize-gcn", "gcn-inference-optimize": "optimize-gcn-inference", "gcn-original": "original-gcn"}


def available():
    return os.path.exists(os.path.join(REF, "algo_kernels", "common_harness", "harness.cpp"))


def build(force=False, verbose=False):
    """Returns {name: path} of the binaries that exist afterwards (prebuilt ones are kept when the reference is absent)."""
    os.makedirs(OUT, exist_ok=True)
    done = {}
    for name, sub in VARIANTS.items():
        exe = os.path.join(OUT, name)
        if not available():
            if os.path.exists(exe):
                done[name] = exe
            continue
        src = os.path.join(REF, "algo_kernels", "common_harness", "harness.cpp")
        deps = [src, os.path.join(REF, "algo_kernels", "vertex_centric", sub, "gcn.h"),
                os.path.join(REF, "include", "ss_vertex_centric_algo_kernel.h"),
                os.path.join(ROOT, "cognn_b200", "host", "shim", "cognn_shim.h"),
                os.path.join(ROOT, "cognn_b200", "host", "shim", "cognn_shim_net.h"),
                os.path.join(ROOT, "cognn_b200", "host", "shim", "include", "cognn_taskworker.h"),
                os.path.join(ROOT, "tests", "ref_standins", "Common", "Defines.h"),
                os.path.join(ROOT, "tests", "ref_standins", "boost", "archive", "binary_oarchive.hpp"), os.path.abspath(__file__)]
        if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
            done[name] = exe
            continue
        # -fwhole-program makes the once-called runAlgoKernelServer a candidate for -finline-functions-called-once; the params lift
        # the caller-size limits that would otherwise veto it (checked after the build: no out-of-line copy may remain)
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fwhole-program", "--param", "large-function-growth=100000",
               "--param", "large-stack-frame-growth=100000", "--param", "large-function-insns=10000000", "--param", "large-unit-insns=10000000",
               "-fopenmp", "-pthread", "-w", "-DSSHEBACKEND", "-DCOGNN_SHIM_IDEAL_NONLINEAR",
               "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "include", "task"),
               "-I" + os.path.join(REF, "algo_kernels", "vertex_centric", sub),
               "-I" + os.path.join(ROOT, "cognn_b200", "host", "shim", "include"), "-I" + os.path.join(ROOT, "tests", "ref_standins"),
               "-I/usr/local/cuda/include", src, "-o", exe, "-L" + os.path.join(ROOT, "cognn_b200"), "-l:libcognn_b200.so",
               "-Wl,--enable-new-dtags,-rpath,$ORIGIN/../../cognn_b200", "-L/usr/local/cuda/lib64", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            if verbose or name != "gcn-original":
                sys.stderr.write(f"[build_ref] {name}: compile failed\n" + r.stderr[-3000:] + "\n")
            if name != "gcn-original":
                raise RuntimeError(f"oracle/_ref/{name} failed to build")
            continue
        sym = subprocess.run("nm -C " + exe + " | grep runAlgoKernelServer | grep -v lambda | grep -c CommSync", shell=True,
                             capture_output=True, text=True).stdout.strip()
        if sym not in ("0", ""):
            sys.stderr.write(f"[build_ref] warning: {name}: runAlgoKernelServer was not inlined; its barrier will dangle\n")
        done[name] = exe
    return done


def build_mock():
    """tests/mock/libcognn_b200.so: the CPU mock of the C-ABI subset the shim calls (oracle-backed), for the CPU drop-in test."""
    sys.path.insert(0, ROOT)
    from oracle import pyoracle

    pyoracle.build()
    src = os.path.join(ROOT, "tests", "mock", "mock_cgb.c")
    lib = os.path.join(ROOT, "tests", "mock", "libcognn_b200.so")
    if os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(src), os.path.getmtime(pyoracle.LIB_PATH)):
        return lib
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", lib, src, "-L" + HERE, "-l:libcgb_oracle.so",
                           "-Wl,--enable-new-dtags,-rpath,$ORIGIN/../../oracle"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_mock())
