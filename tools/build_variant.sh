#!/bin/bash
# Builds an ablation/variant copy of libsketchdist.so with extra -D flags:  tools/build_variant.sh TAG -DSKD_X=1 ...
# Output: tools/variants/libskd_TAG.so (select it with SKD_LIB_OVERRIDE).  Not part of the product build.
set -e
tag=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$here/tools/variants"
python - "$here" "$tag" "$@" <<'PY'
import importlib.util, sys
here, tag, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
spec = importlib.util.spec_from_file_location("skd_build", here + "/sketchlib.rust_b200/build.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(m.build(extra_flags=flags, out=f"{here}/tools/variants/libskd_{tag}.so"))
PY
