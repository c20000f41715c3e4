"""Module-level configuration, same names and meanings as the reference's spnet/config.py
(read at call time, e.g. cf.loss_type when custom_loss runs — spnet/models.py:569)."""
import numpy as np

dtype = np.float32
meta_extension = ".csv"

# colours (BGR, for the optional PNG drawing)
blue = (255, 0, 0)
red = (0, 0, 255)
green = (0, 200, 0)
white = (255, 255, 255)
black = (0, 0, 0)
grey = (128, 128, 128)
lightgrey = (210, 210, 210)
yellow = (255, 255, 0)[::-1]
cyan = (0, 220, 220)[::-1]
veridis_purple = (72, 18, 84)[::-1]
truecolor = yellow
predcolor = veridis_purple

# column layout of one predictor (spnet/config.py:30-38)
vars_per_pred = 8
ind_cx = 0
ind_cy = 1
ind_semi_a = 2
ind_semi_b = 3
ind_angle1 = 4   # cos(2 theta)
ind_angle2 = 5   # sin(2 theta)
ind_noobj = 6
ind_rings = 7

loss_type = "same"        # 'same' = MSE everywhere; anything else = BCE-with-logits on noobj
model_type = "monolithic"  # 'big' = no resize (384x512); default resizes to 331x331
basemodel = "Xception"

# B200 engine options (not in the reference)
compute_dtype = "bf16"    # 'bf16' | 'fp32'
