"""Self-contained command line equal to the reference driver VBx/vbhmm.py:55-179, with the same options, but every
recording of the archive is processed in ONE batch on the GPU (front end, AHC, VB-HMM, labels):

    python -m vbx_b200.cli --init AHC+VB --out-rttm-dir exp --xvec-ark-file exp/ES2005a.ark \\
        --segments-file exp/ES2005a.seg --xvec-transform VBx/models/ResNet101_16kHz/transform.h5 \\
        --plda-file VBx/models/ResNet101_16kHz/plda --threshold -0.015 --lda-dim 128 --Fa 0.3 --Fb 17 --loopP 0.99

Reads Kaldi ark / segments / PLDA (binary or text) / transform.h5 through vbx_b200.formats (no kaldi_io, h5py or
fastcluster needed) and writes one RTTM per recording, formatted as VBx/vbhmm.py:48-51.
"""
import argparse
import os
import sys

import numpy as np


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    # option names, types and defaults of VBx/vbhmm.py:55-102
    ap.add_argument('--init', required=True, type=str, choices=['AHC', 'AHC+VB'])
    ap.add_argument('--out-rttm-dir', required=True, type=str)
    ap.add_argument('--xvec-ark-file', required=True, type=str)
    ap.add_argument('--segments-file', required=True, type=str)
    ap.add_argument('--xvec-transform', required=True, type=str)
    ap.add_argument('--plda-file', required=True, type=str)
    ap.add_argument('--threshold', required=True, type=float)
    ap.add_argument('--lda-dim', required=True, type=int)
    ap.add_argument('--Fa', required=True, type=float)
    ap.add_argument('--Fb', required=True, type=float)
    ap.add_argument('--loopP', required=True, type=float)
    ap.add_argument('--target-energy', required=False, type=float, default=1.0)     # unused by the cosine AHC, as in the reference
    ap.add_argument('--init-smoothing', required=False, type=float, default=5.0)
    ap.add_argument('--output-2nd', required=False, type=bool, default=False)
    # additions
    ap.add_argument('--chain', default='auto', choices=['auto', 'tcgen05', 'float64'],
                    help='front end arithmetic: fused tensor-core kernels (float32-level) or float64 torch ops')
    ap.add_argument('--device', default=None, help='CUDA device, e.g. cuda:0 (default: the current device)')
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    assert 0 <= args.loopP <= 1, f'Expecting loopP between 0 and 1, got {args.loopP} instead.'     # VBx/vbhmm.py:103
    from . import formats
    from .pipeline import diarize_batch
    segs = formats.read_segments(args.segments_file)                        # VBx/vbhmm.py:105
    plda = formats.read_kaldi_plda(args.plda_file)                          # VBx/vbhmm.py:107
    mean1, mean2, lda = formats.read_xvec_transform(args.xvec_transform)    # VBx/vbhmm.py:125-128
    recs = {}
    for name, (keys, x) in formats.read_xvectors_by_recording(args.xvec_ark_file).items():     # VBx/vbhmm.py:117-123
        print(name)
        seg_names, times = segs[name]
        assert np.all(np.array(seg_names) == np.array(keys))               # VBx/vbhmm.py:166
        recs[name] = (x, times)
    out = diarize_batch(recs, (mean1, mean2, lda), plda, Fa=args.Fa, Fb=args.Fb, loopP=args.loopP, lda_dim=args.lda_dim,
                        threshold=args.threshold, smoothing=args.init_smoothing, init=args.init, chain=args.chain,
                        device=args.device, output_2nd=args.output_2nd)
    os.makedirs(args.out_rttm_dir, exist_ok=True)                           # VBx/vbhmm.py:170
    for name, item in out.items():
        with open(os.path.join(args.out_rttm_dir, f'{name}.rttm'), 'w') as fp:
            fp.write(''.join(line + os.linesep for line in item['rttm']))
        if item['rttm2nd'] is not None:
            d2 = f'{args.out_rttm_dir}2nd'
            os.makedirs(d2, exist_ok=True)
            with open(os.path.join(d2, f'{name}.rttm'), 'w') as fp:
                fp.write(''.join(line + os.linesep for line in item['rttm2nd']))
    return 0


if __name__ == '__main__':
    sys.exit(main())
