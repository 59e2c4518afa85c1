"""One streaming-kernel variant (argv[1]) on a 4096^2 CMOS frame, for an ncu --set full capture."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from microbench import engine_for
from scopyon_b200 import _native
lib = _native.load()
ADC = "analog_to_digital_converter: {type: column, count: 2.0, offset: 100, fullwell: 30000}"
size = 4096
configs, eng = engine_for("default:\n    detector: {type: CMOS, image_size: [%d, %d], QE: 0.73}\n    %s\n" % (size, size, ADC))
gen = torch.Generator(device="cuda").manual_seed(5)
photons = torch.empty((size, size), dtype=torch.float32, device="cuda").exponential_(1.0 / 0.6, generator=gen)
adc = torch.empty_like(photons)
if hasattr(lib, "scb_detector_set_variant"):
    lib.scb_detector_set_variant(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
for _ in range(3):
    eng.detect(photons, 1, 42, adc=adc)
torch.cuda.synchronize()
