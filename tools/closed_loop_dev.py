import torch, sys
sys.path.insert(0, ".")
import ivclab_b200 as ivc
from bench_configs import luma_seq, timed
s5 = luma_seq(60, 1080, 1920, 5000)
pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="exact")
rec = s5[:1] + 0.25
t = timed(lambda: pc.estimate(rec, s5[1:2]), 20)
cl = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact", use_graph=False)
t2 = timed(lambda: cl.code_sequence(s5), 3, warm=1)
q = luma_seq(21, 144, 176, 2)
t3 = timed(lambda: cl.code_sequence(q), 10)
print(f"exact ME one 1080p frame {t*1e3:.1f} us; closed loop {t2/60*1e3:.1f} us/frame; QCIF closed loop {t3/21*1e3:.1f} us/frame")
