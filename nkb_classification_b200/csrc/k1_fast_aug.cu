// The TMA-staged K1 kernel with the train-time augmentations fused (HorizontalFlip, VerticalFlip,
// RandomBrightnessContrast, HueSaturationValue, CoarseDropout): the same source as k1_fast.cu, instantiated with AUG.
#define K1_FAST_AUG true
#define K1_FAST_LAUNCHER launch_k1_fast_aug
#include "k1_fast.cu"
