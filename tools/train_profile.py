import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200.training import TrainableQFormer
from mraudio_b200.xinstructblip import XInstructBLIPQFormers
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video",)).to(dev)
model.freeze_qformers(False)
st = TrainableQFormer(model.video_Qformer, model.video_query_tokens, model.video_llm_proj)
g = torch.Generator().manual_seed(1)
rows, T = 64, 32
enc = torch.randn(rows, 257, 1408, generator=g).to(torch.bfloat16).to(dev)
ids = torch.randint(1000, 30000, (rows, T), generator=g).to(dev)
atts = torch.ones(rows, 32 + T, dtype=torch.long, device=dev)
G = torch.randn(rows, 32, 4096, generator=g).to(dev)
def ev(): 
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    e0 = ev(); y = st.forward(enc, ids, atts); e1 = ev()
    loss = (y.float() * G).sum(); e2 = ev()
    loss.backward(); e3 = ev()
    st.adam_step(1e-4, zero_grad=True); e4 = ev(); e5 = ev()
    torch.cuda.synchronize()
    print(f"fwd {e0.elapsed_time(e1):.2f} | loss {e1.elapsed_time(e2):.2f} | bwd {e2.elapsed_time(e3):.2f} | adam+refresh {e3.elapsed_time(e4):.2f} | zero {e4.elapsed_time(e5):.2f}  (bwd launches {st.last_backward_launches})")
