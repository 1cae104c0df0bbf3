import sys; sys.path.insert(0, "/root/repo")
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")
from myrtlespeech_b200 import _lib
lib = _lib.load()
for k in (b"max_ctas_fwd_c2", b"max_ctas_fwd_c4", b"max_ctas_mega_c2", b"max_ctas_mega_c4"):
    print(k.decode(), lib.rnnt_debug_get(k))
