"""Batch Dice evaluation (drop-in for reference segmentation3d/core/seg_eval.py:8-57): the caller on the far side of the
inference path - it reads the masks `segmentation()` wrote and scores them with utils/metrics.py::cal_dsc, the metric
the parity bars are stated in.  Host-side glue (file reads + a csv), same csv layout as the reference:
one row per case, `label<k>_score` / `label<k>_type` columns, then a `mean` and a `std` row."""
import os

import pandas as pd

from segmentation3d.utils.image3d import read_image
from segmentation3d.utils.metrics import cal_dsc


def cal_dsc_batch(gt_files, seg_files, labels, threshold, save_csv_file_path):
    """gt_files / seg_files: parallel lists of mask paths; labels: which labels to score; threshold: a label with fewer
    voxels than this counts as absent (TN / FP / FN instead of a Dice value); the table is written to
    save_csv_file_path and returned."""
    assert isinstance(gt_files, list) and isinstance(seg_files, list)
    assert len(gt_files) == len(seg_files)
    rows = []
    for gt_path, seg_path in zip(gt_files, seg_files):
        gt, seg = read_image(gt_path).to_numpy(), read_image(seg_path).to_numpy()
        case_name = os.path.basename(gt_path)                      # the reference's second assignment wins (:32-33)
        row = [case_name]
        for label in labels:
            score, kind = cal_dsc(gt, seg, label, threshold)
            row.extend([score, kind])
            print('case_name: {}, label: {}, score: {}, type: {}'.format(case_name, label, score, kind))
        rows.append(row)
    columns = ['filename']
    for label in labels:
        columns.extend(['label{}_score'.format(label), 'label{}_type'.format(label)])
    df = pd.DataFrame(data=rows, columns=columns)
    stats = [['mean'], ['std']]
    for label in labels:
        col = df['label{}_score'.format(label)]
        mean, std = col.mean(), col.std()
        print(mean, std)
        stats[0].extend([mean, 'ignore_type'])
        stats[1].extend([std, 'ignore_type'])
    df = pd.concat([df, pd.DataFrame(data=stats, columns=columns)])   # DataFrame.append (reference :56) left pandas in 2.0
    df.to_csv(save_csv_file_path)
    return df
