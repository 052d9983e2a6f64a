"""`seg_infer` console script (drop-in for reference segmentation3d/seg_infer.py:6-52):
-i/--input, -m/--model, -o/--output, -n/--seg_name, -g/--gpu_id, --save_image, --save_prob."""
import argparse

from segmentation3d.core.seg_infer import segmentation


def main():
    parser = argparse.ArgumentParser(
        description='Inference engine for 3d medical image segmentation (B200 build). Input: a single image, '
                    'a text file listing test images, or a folder of images.')
    parser.add_argument('-i', '--input', required=True, help='input folder/file for intensity images')
    parser.add_argument('-m', '--model', required=True, help='model root folder')
    parser.add_argument('-o', '--output', required=True, help='output folder for segmentation')
    parser.add_argument('-n', '--seg_name', default='seg.mha', help='the name of the segmentation result to be saved')
    parser.add_argument('-g', '--gpu_id', type=int, default=0, help='the gpu id to run model (this build has no CPU path)')
    parser.add_argument('--save_image', action='store_true', help='whether to save original image')
    parser.add_argument('--save_prob', action='store_true', help='whether to save all prob maps')
    args = parser.parse_args()
    segmentation(args.input, args.model, args.output, args.seg_name, args.gpu_id, False, True, args.save_image, args.save_prob)


if __name__ == '__main__':
    main()
