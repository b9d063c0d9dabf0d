"""Build libcdml.so (sm_100a only) in-tree with nvcc.  `python -m cdml_b200.build` or __graft_entry__.build()."""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcdml.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found; libcdml.so cannot be built")


def _sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
  h = hashlib.sha256()
  for root in (CSRC, os.path.join(HERE, "..", "include")):
    for f in sorted(os.listdir(root)):
      if f.endswith((".cu", ".cuh", ".h")):
        with open(os.path.join(root, f), "rb") as fh:
          h.update(f.encode())
          h.update(fh.read())
  h.update(" ".join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force=False, verbose=False):
  """Compile every .cu under csrc/ for sm_100a and link libcdml.so next to this file."""
  stamp = os.path.join(OBJ_DIR, "stamp")
  digest = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
    return LIB
  os.makedirs(OBJ_DIR, exist_ok=True)
  nvcc = _nvcc()

  def compile_one(src):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
      cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
      sys.stderr.write(r.stderr)
    return obj

  with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    objs = list(ex.map(compile_one, _sources()))
  r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                     capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
  with open(stamp, "w") as f:
    f.write(digest)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
