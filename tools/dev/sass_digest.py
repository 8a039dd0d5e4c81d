"""md5 of the SASS instruction listing of libgsr_b200.so (comments / line tables do not enter it).

    python tools/dev/sass_digest.py [path/to/libgsr_b200.so]

Used to show that the library built from the committed sources is, instruction for instruction, the one a GPU run
measured (profiles/README.md names the digest of the round's last GPU call)."""
import hashlib
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parents[2] / "gaussian_splatting_b200" / "libgsr_b200.so"
text = subprocess.run(["cuobjdump", "-sass", str(lib)], check=True, capture_output=True, text=True).stdout
lines = [l + "\n" for l in text.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
print(hashlib.md5("".join(lines).encode()).hexdigest(), len(lines), "instructions", lib)
