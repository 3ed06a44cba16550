"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path,
packed where they lie under /root/reference into one import archive, so that the reference arm of bench.py
(`--impl reference`, `cpu_baseline.kind = "reference"`) times the stock code path on the GPU box, where
/root/reference does not exist.

    python oracle/build_ref.py          # -> oracle/_ref/cgat_reference.zip   (git-ignored, travels with gpurun)

Nothing is copied into the repository's history: oracle/_ref/ is listed in .gitignore (like the built .so files).
The archive holds CGAT/{__init__,CGAT,roost_message,message_changed,Hypernetworksmp}.py byte for byte; Python imports
straight from it (zipimport) once the stand-ins of oracle/standins.py for torch_scatter / torch_geometric (not
installable here) are registered.  Only bench.py's reference arm / cpu_baseline and tests may use it;
nothing under cgat_b200/ does."""
from __future__ import annotations

import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ZIP = os.path.join(HERE, "_ref", "cgat_reference.zip")
FILES = ("__init__.py", "CGAT.py", "roost_message.py", "message_changed.py", "Hypernetworksmp.py")


def build_ref(reference_root="/root/reference"):
    """Write oracle/_ref/cgat_reference.zip from `reference_root` if it exists; return the archive path or None."""
    src = os.path.join(reference_root, "CGAT")
    if not all(os.path.exists(os.path.join(src, f)) for f in FILES):
        return REF_ZIP if os.path.exists(REF_ZIP) else None
    os.makedirs(os.path.dirname(REF_ZIP), exist_ok=True)
    tmp = REF_ZIP + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            info = zipfile.ZipInfo("CGAT/" + f, date_time=(2020, 1, 1, 0, 0, 0))   # deterministic archive
            with open(os.path.join(src, f), "rb") as fh:
                z.writestr(info, fh.read(), zipfile.ZIP_DEFLATED)
    os.replace(tmp, REF_ZIP)
    return REF_ZIP


def import_ref():
    """The unmodified reference modules from the archive: {'CGAT': module, 'roost_message': ..., ...} or None."""
    if not os.path.exists(REF_ZIP):
        return None
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import standins
    return standins.import_reference(REF_ZIP)


if __name__ == "__main__":
    print(build_ref())
