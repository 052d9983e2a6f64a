"""Training configuration template (same keys as reference config/train_config.py:5-140)."""
from easydict import EasyDict as edict
from segmentation3d.utils.normalizer import FixedNormalizer, AdaptiveNormalizer

__C = edict()
cfg = __C

__C.general = {}
__C.general.imseg_list = '/path/to/train.txt'      # txt: N, then image/mask path pairs; or csv with image_path,mask_path
__C.general.save_dir = '/path/to/model_folder'
__C.general.model_scale = 'fine'
__C.general.resume_epoch = -1                      # -1: from scratch
__C.general.num_gpus = 1                           # >1: launch one process per GPU with torchrun
__C.general.seed = 0

__C.dataset = {}
__C.dataset.num_classes = 2
__C.dataset.spacing = [1.0, 1.0, 1.0]
__C.dataset.crop_size = [96, 96, 96]               # multiples of max_stride = 16
__C.dataset.sampling_method = 'HYBRID'             # CENTER | GLOBAL | MASK | HYBRID
__C.dataset.interpolation = 'LINEAR'
__C.dataset.crop_normalizers = [AdaptiveNormalizer()]
__C.dataset.random_translation = [15, 15, 15]      # mm
__C.dataset.random_scale = [0.9, 1.1]

__C.loss = {}
__C.loss.name = 'Dice'                             # Focal | Dice | CE
__C.loss.obj_weight = [1 / 2, 1 / 2]
__C.loss.focal_gamma = 2

__C.net = {}
__C.net.name = 'vnet'                              # module under segmentation3d.network

__C.train = {}
__C.train.epochs = 1001
__C.train.batchsize = 8
__C.train.num_threads = 4
__C.train.lr = 1e-4
__C.train.betas = (0.9, 0.999)
__C.train.save_epochs = 100

__C.debug = {}
__C.debug.save_inputs = False
