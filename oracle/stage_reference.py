"""Stage the read-only reference checkout where the GPU box can see it.  TEST / MEASUREMENT INFRASTRUCTURE.

``baseline/_ref`` is git-ignored (it never enters history) but travels with the gpurun snapshot, so that on the
GPU box

* ``bench.py`` times the VERBATIM reference on the host cores (``cpu_baseline.kind = "reference"``), and
* ``tests/test_gpu_train_dropin.py`` runs the reference's own ``train()`` loop on the GPU.

The reference is plain Python without a ``setup.py``, so the base recipe
(``pip install --target baseline/_ref /root/reference``) has nothing to build: the install is a copy of its source
tree.  Nothing in ``cgc_net_b200/`` reads it.

    python oracle/stage_reference.py [--src /root/reference] [--remove]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
ITEMS = ('model', 'common', 'dataflow', 'train.py', 'setting.py', 'requirements.txt')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    ap.add_argument('--remove', action='store_true', help='delete the staged copy')
    a = ap.parse_args()
    if a.remove:
        shutil.rmtree(DST, ignore_errors=True)
        print('removed', DST)
        return
    if not os.path.isfile(os.path.join(a.src, 'model', 'network.py')):
        raise SystemExit('no reference checkout at %s' % a.src)
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        src, dst = os.path.join(a.src, item), os.path.join(DST, item)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__'))
        elif os.path.isfile(src):
            shutil.copy(src, dst)
    print('staged %s -> %s' % (a.src, DST))


if __name__ == '__main__':
    main()
