#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified* reference kernel layer
# (/root/reference/velocyto/speedboosted.pyx, the only native component of
# velocyto.py) into oracle/_ref/ so that the restated oracle in this directory
# can be pinned against the reference itself, and so that bench.py can time the
# reference's own CPU code path (cpu_baseline.kind == "reference").
#
# Nothing from /root/reference is copied into the repository: the .pyx is read
# where it lies, the Cython-generated C and the .so land in oracle/_ref/, which
# is git-ignored (but not gpurun-ignored, so the built .so travels to the GPU
# box).  Flags are the reference's own: -fopenmp -ffast-math (setup.py:17-21).
# /usr/bin/gcc is forced because other gcc installs lack libgomp.spec.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${VELO_REFERENCE_ROOT:-/root/reference}"
pyx="$ref/velocyto/speedboosted.pyx"
out="$here/_ref"
if [ ! -f "$pyx" ]; then
  echo "build_ref: $pyx not present (GPU box?) - keeping prebuilt files in $out" >&2
  exit 0
fi
mkdir -p "$out"
py="${PYTHON:-python}"
suffix="$($py -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
pyinc="$($py -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
npinc="$($py -c 'import numpy; print(numpy.get_include())')"
target="$out/speedboosted$suffix"
if [ -f "$target" ] && [ "$target" -nt "$pyx" ]; then
  exit 0
fi
"$py" -m cython -3 "$pyx" -o "$out/speedboosted.c"
/usr/bin/gcc -shared -fPIC -O2 -fopenmp -ffast-math \
  -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION \
  -I"$pyinc" -I"$npinc" "$out/speedboosted.c" -o "$target"
echo "build_ref: built $target"
