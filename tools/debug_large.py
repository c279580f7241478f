"""Debug helper: B=1024 uint8 images (18.5 GB slab), compare every engine against torch slicing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jolineedle_b200.gather import ImageSet
b, P, gh, gw = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 448, 5, 6
g = torch.Generator(device="cuda").manual_seed(5)
base = torch.randint(0, 256, (24, 3, gh * P, gw * P), dtype=torch.uint8, device="cuda", generator=g)
images = base.repeat(b // 24 + 1, 1, 1, 1)[:b].contiguous()
print("slab bytes", images.numel(), "ptr", hex(images.data_ptr()))
table = (torch.arange(256, dtype=torch.uint8).float() / 255).cuda()
s = ImageSet(images, P)
pos = torch.stack([torch.randint(0, gh, (b,), device="cuda", generator=g), torch.randint(0, gw, (b,), device="cuda", generator=g)], 1)
for normalize in (True, False):
    for engine in ("tensor", "bulk", "ldg"):
        out = s.gather(pos, normalize=normalize, engine=engine)
        torch.cuda.synchronize()
        bad = []
        for i in range(b):
            y, x = pos[i].tolist()
            tile = images[i, :, y * P:(y + 1) * P, x * P:(x + 1) * P]
            want = table[tile.long()] if normalize else tile
            if not torch.equal(out[i], want):
                d = (out[i] != want)
                rows = d.any(dim=2).nonzero()
                bad.append((i, int(d.sum()), rows[0].tolist(), rows[-1].tolist()))
        print(f"normalize={normalize} engine={engine}: {len(bad)} bad tiles of {b}; first: {bad[:6]}")
