"""Generates tests/golden/gtm_demo_digest.json from the reference's demo streams (docs/demo/*.gtm): header fields and a
SHA-256 of the first frames as decoded by tiler_b200.gtm.  Run in the build container (needs /root/reference)."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tiler_b200 import gtm
out = {}
for name in ("city_cif.gtm", "football_cif.gtm"):
    d = open("/root/reference/docs/demo/" + name, "rb").read()
    hdr = gtm.parse_header(d)
    frames, _ = gtm.decode_gtm(d, max_frames=48)
    out[name] = {"header": [hdr[k] for k in ("width", "height", "kf_count", "frame_count")], "frames": 48,
                 "sha256": hashlib.sha256(frames.tobytes()).hexdigest()}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gtm_demo_digest.json"), "w"), indent=1)
print(out)
