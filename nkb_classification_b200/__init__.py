"""nkb_classification_b200 -- the B200-native hot path of nkb-classification.

Per-sample crop / cv2-exact resize / normalize / layout (K1), the multitask
heads with fused loss and gradients (K2), argmax + confusion counts (K3) and the
batch-sharded all-reduce (K4), behind the reference's config-driven API:

    dataset.get_dataset / get_inference_dataset     model.get_model
    losses.get_loss                                  engine.train_epoch / val_epoch
    logging.BaseLogger                               metrics.compute_metrics
    inference.inference                              utils.get_optimizer / get_scheduler

All arithmetic on the path runs in libnkbk.so (include/nkbk.h); there is no CPU
or PyTorch fallback -- importing works anywhere, calling needs the built library
and a CUDA device.
"""
__version__ = "0.1.0"
