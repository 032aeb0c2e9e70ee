"""Generates tests/golden/hard_eval_golden.pt: eval-mode runs of the UNMODIFIED reference with `hard_mode: True`
(min instead of sum in the quantifiers when answers are given, batch_base_types.py:104-112) on the inputs and weights of
every golden_*_s1.pt fixture.  Run here (build container):  python tests/golden/make_golden_hard.py"""

import glob
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from ref_harness import ReferenceRun  # noqa: E402


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(HERE, 'golden_*_s1.pt'))):
        case = torch.load(path, weights_only=False)
        run = ReferenceRun(case['metadata'], case['dims'], seed=0, config_overrides={'hard_mode': True})
        # the reference registers the same tensors under many `_ops.*` paths: every alias gets the fixture's value
        full = run.state_dict()
        for k in list(full):
            for name, v in case['state'].items():
                if k == name or k.endswith('.' + name.split('.', 1)[1]) and k.split('.')[-4:] == name.split('.')[-4:] \
                        and name.split('.')[1] in k:
                    full[k] = v
        run.load_state_dict(full)
        check = run.state_dict()
        assert all(torch.equal(check[k], v) for k, v in case['state'].items())
        pbs = run.collate(json.loads(case['questions']), case['features'], case['batch_index'], split_num=1)
        ev = run.forward(pbs, is_training=False)
        rec = {'log_probability': ev['log_probability'].detach().clone(), 'answer': ev['answer'], 'type': int(ev['type'])}
        if int(ev['type']) == 1:
            rec['options'] = [list(o) for o in ev['options']]
        out[os.path.basename(path)] = rec
        print(os.path.basename(path), rec['log_probability'].numel())
    torch.save(out, os.path.join(HERE, 'hard_eval_golden.pt'))


if __name__ == '__main__':
    main()
