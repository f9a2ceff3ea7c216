import time, numpy as np, sys
sys.path.insert(0, '.')
from libnativecpurenderer_b200.binding import Renderer
R = Renderer()
mask = R.Texture.from_numpy(np.full((512, 512, 4), 255, np.uint8))
t0 = time.time()
tex = [R.lib.CreateMilthmHitEffectTexture(mask._ptr, 0.25, k / 16.0, 150 / 255, 144 / 255, 253 / 255) for k in range(16)]
print("hit-effect 512^2 textures: %.1f ms each" % ((time.time() - t0) / 16 * 1e3))
