"""Device time of the cluster QRCP at the config-4 sizes for the cluster size given by ENLSIP_QC_NC (read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_check_small import time_qrcp
print("ENLSIP_QC_NC=%s" % os.environ.get("ENLSIP_QC_NC", "auto"),
      " ".join("%dx%d: %.3f ms" % (r, c, time_qrcp(r, c, reps=3) * 1e3) for r, c in ((256, 64), (64, 64), (257, 192), (513, 384))))
