import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jolineedle_b200.gather import ImageSet
from oracle.gaze_oracle import translate_oracle
case = sys.argv[1]
P, gh, gw, b = 64, 4, 5, 2
g = torch.Generator().manual_seed(0)
images = torch.rand((b, 3, gh * P, gw * P), generator=g)
shifts = {"zero": [(0, 0), (0, 0)], "pos": [(16, 8), (4, 4)], "odd": [(13, 7), (1, 3)], "neg": [(-13, -7), (-64, -1)],
          "part": [(300, 0), (0, 250)], "full": [(400, 0), (0, 300)]}[case]
s = ImageSet(images.cuda(), P)
pos = torch.tensor([[0, 0], [3, 4]], dtype=torch.int64).cuda()
d = torch.tensor([(ty, tx) for (tx, ty) in shifts], dtype=torch.int32).cuda()
got = s.gather(pos, shifts=d, engine=sys.argv[2] if len(sys.argv) > 2 else "tensor")
torch.cuda.synchronize()
sh = translate_oracle(images, shifts)
want = torch.stack([sh[i][:, y * P:(y + 1) * P, x * P:(x + 1) * P] for i, (y, x) in enumerate(pos.tolist())])
print(case, "equal:", torch.equal(got.cpu(), want))
