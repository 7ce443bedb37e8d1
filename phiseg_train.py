#!/usr/bin/env python
"""Thin training CLI with the reference's interface (phiseg_train.py:16-50): `python phiseg_train.py EXP_PATH`.

Loads the experiment module by path (reference files load unchanged, phiseg/experiments/__init__.py), makes
<log_root>/<log_dir_name>/<experiment_name>, copies the experiment file there (the evaluation scripts of the reference glob
that copy, phiseg_test_quantitative.py:93-97), builds the data object and the model, and calls phiseg.train(data).

Data (data/data_switch.py -> data/lidc_data.py of the reference): `--data FILE.npz` reads arrays train_images / train_labels
/ val_images / val_labels [/ test_*] (images [N,H,W] float, labels [N,H,W,A] uint8: what lidc_data_loader.py writes into
its HDF5 file; an .hdf5 / .h5 file with the reference's dataset names is read when h5py is importable).  `--synthetic N`
trains on N LIDC-shaped synthetic images instead (the dataset itself cannot be shipped).  The batches are produced on the
device by data.BatchProvider (augmentation included, one launch per batch).

    python phiseg_train.py phiseg-code_b200/phiseg/experiments/phiseg_7_5.py --synthetic 256 --num-iter 200
"""
import argparse
import logging
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')


def load_arrays(path):
    """{'train' | 'val' | 'test': {'images', 'labels'}} from an .npz or (with h5py) the reference's preprocessed HDF5 file
    (data/lidc_data_loader.py:92-131: datasets images_train, labels_train, images_val, ...)."""
    out = {}
    if path.endswith('.npz'):
        f = np.load(path)
        for tt in ('train', 'val', 'test'):
            if tt + '_images' in f.files:
                out[tt] = {'images': f[tt + '_images'], 'labels': f[tt + '_labels']}
    else:
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError('reading %s needs h5py, which is not installed here: convert the file to .npz '
                               '(train_images, train_labels, val_images, val_labels)' % path) from e
        f = h5py.File(path, 'r')
        for tt in ('train', 'val', 'test'):
            if 'images_' + tt in f:
                out[tt] = {'images': f['images_' + tt][()], 'labels': f['labels_' + tt][()]}
    if 'train' not in out or 'val' not in out:
        raise RuntimeError('%s holds no train / val arrays' % path)
    return out


def prepare(exp_path, log_root=None):
    """phiseg_train.py:37-47: the experiment module, its log directory, and the copy of the experiment file inside it."""
    from __graft_entry__ import load_package
    load_package()
    import importlib
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp_config = ex.load_experiment(exp_path)
    root = log_root or getattr(exp_config, 'log_root', None) or os.environ.get('PHISEG_LOG_ROOT', './logs')
    exp_config.log_root = root
    log_dir = os.path.join(root, exp_config.log_dir_name, exp_config.experiment_name)
    os.makedirs(log_dir, exist_ok=True)
    dst = os.path.join(log_dir, os.path.basename(exp_path))
    if os.path.abspath(dst) != os.path.abspath(exp_path):
        shutil.copy(exp_path, dst)
    logging.info('!!!! Copied exp_config file to experiment folder !!!!')
    return exp_config, log_dir


def main(exp_config, data_path=None, synthetic=0, mode=None):
    """phiseg_train.py:16-30"""
    import importlib
    logging.info('**************************************************************')
    logging.info(' *** Running Experiment: %s', exp_config.experiment_name)
    logging.info('**************************************************************')
    D = importlib.import_module('phiseg_code_b200.data')
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    if synthetic:
        syn = D.SyntheticLIDC(num_train=synthetic, num_val=max(8, synthetic // 8), size=exp_config.image_size[0],
                              nlabels=exp_config.nlabels, annotators=exp_config.num_labels_per_subject)
        arrays = {'train': {'images': syn.train.images, 'labels': syn.train.labels},
                  'val': {'images': syn.validation.images, 'labels': syn.validation.labels}}
    elif data_path:
        arrays = load_arrays(data_path)
    else:
        raise SystemExit('no data: pass --data FILE.npz (or .hdf5) or --synthetic N; the reference reads %s'
                         % getattr(exp_config, 'preproc_folder', '<preproc_folder>'))
    model = pm.phiseg(exp_config, mode=mode)
    data = D.lidc_data(exp_config, arrays, device=model.device)
    model.train(data)
    return model


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Script for training')
    parser.add_argument('EXP_PATH', type=str, help='Path to experiment config file')
    parser.add_argument('--data', type=str, default=None, help='preprocessed data (.npz, or the reference HDF5 with h5py)')
    parser.add_argument('--synthetic', type=int, default=0, help='train on N synthetic LIDC-shaped images')
    parser.add_argument('--num-iter', type=int, default=None, help='override exp_config.num_iter')
    parser.add_argument('--log-root', type=str, default=None, help='default: $PHISEG_LOG_ROOT or ./logs')
    parser.add_argument('--mode', type=str, default=None, choices=['fast', 'parity_tc', 'parity'])
    return parser.parse_args(argv)


if __name__ == '__main__':
    args = parse_args()
    exp_config, log_dir = prepare(args.EXP_PATH, args.log_root)
    if args.num_iter is not None:
        exp_config.num_iter = args.num_iter
    main(exp_config, data_path=args.data, synthetic=args.synthetic, mode=args.mode)
