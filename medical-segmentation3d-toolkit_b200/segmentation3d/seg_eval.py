"""`seg_eval` console script.  The reference's segmentation3d/seg_eval.py:7-33 hard-codes its paths; the same flow
(case list -> <gt_folder>/<case>/<gt_name> vs <seg_folder>/<case>/<seg_name> -> csv) with the paths as flags."""
import argparse
import os

from segmentation3d.core.seg_eval import cal_dsc_batch
from segmentation3d.core.seg_infer import read_test_csv, read_test_txt


def main():
    parser = argparse.ArgumentParser(description='Dice evaluation of segmentation results against ground truth masks.')
    parser.add_argument('-i', '--input', required=True, help='test list (.txt or .csv) naming the cases')
    parser.add_argument('--gt_folder', required=True)
    parser.add_argument('--gt_name', default='seg.mha')
    parser.add_argument('--seg_folder', required=True)
    parser.add_argument('--seg_name', default='seg.mha')
    parser.add_argument('-l', '--labels', type=int, nargs='+', default=[1])
    parser.add_argument('-t', '--threshold', type=int, default=10)
    parser.add_argument('-o', '--output', required=True, help='result csv')
    args = parser.parse_args()
    if args.input.endswith('.txt'):
        case_list, _ = read_test_txt(args.input)
    elif args.input.endswith('.csv'):
        case_list, _ = read_test_csv(args.input)
    else:
        raise ValueError('Unsupported file')
    gt_files = [os.path.join(args.gt_folder, c, args.gt_name) for c in case_list]
    seg_files = [os.path.join(args.seg_folder, c, args.seg_name) for c in case_list]
    cal_dsc_batch(gt_files, seg_files, args.labels, args.threshold, args.output)


if __name__ == '__main__':
    main()
