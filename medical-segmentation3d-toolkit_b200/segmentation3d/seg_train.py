"""`seg_train` console script (drop-in for reference segmentation3d/seg_train.py:6-18): -i/--input <train_config.py>."""
import argparse

from segmentation3d.core.seg_train import train


def main():
    parser = argparse.ArgumentParser(description='Training engine for 3d medical image segmentation (B200 build)')
    parser.add_argument('-i', '--input', required=True, help='training config file')
    args = parser.parse_args()
    train(args.input)


if __name__ == '__main__':
    main()
