"""Training configuration template.  `seg_train -i <this file>` reads the module-level `cfg`
(key-for-key compatible with the reference's config/train_config.py)."""
from easydict import EasyDict as edict
from segmentation3d.utils.normalizer import FixedNormalizer, AdaptiveNormalizer  # noqa: F401  (both usable below)

cfg = edict(dict(
    general=dict(
        imseg_list='/path/to/train.txt',   # txt: N, then image/mask path pairs - or csv with image_path,mask_path columns
        save_dir='/path/to/model_folder',  # <save_dir>/<model_scale>/checkpoints/chk_<epoch>/ is written here
        model_scale='fine',
        resume_epoch=-1,                   # -1: train from scratch (the model_scale folder is wiped)
        num_gpus=1,                        # >1: launch one process per GPU with torchrun
        seed=0,
    ),
    dataset=dict(
        num_classes=2,
        spacing=[1.0, 1.0, 1.0],           # mm; training images must already be at this spacing in this build
        crop_size=[96, 96, 96],            # voxels, multiples of the network's max_stride (16)
        sampling_method='HYBRID',          # CENTER | GLOBAL | MASK | HYBRID
        interpolation='LINEAR',            # LINEAR | NN, stored in the checkpoint for inference-time resampling
        random_translation=[15, 15, 15],   # mm
        random_scale=[0.9, 1.1],
    ),
    loss=dict(name='Dice', obj_weight=[1 / 2, 1 / 2], focal_gamma=2),      # name: Focal | Dice | CE
    net=dict(name='vnet'),                                                 # a module under segmentation3d.network
    train=dict(epochs=1001, batchsize=8, num_threads=4, lr=1e-4, betas=(0.9, 0.999), save_epochs=100),
    debug=dict(save_inputs=False),
))
# normaliser objects are stored by reference (they are serialised into params.pth through to_dict())
cfg.dataset.crop_normalizers = [AdaptiveNormalizer()]
