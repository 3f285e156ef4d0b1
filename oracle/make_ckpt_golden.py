"""Generate tests/golden_eval/reference_checkpoint_mf.pth with the reference's own model class and checkpoint dict (TEST INFRASTRUCTURE).

    python oracle/make_ckpt_golden.py        # build container only (/root/reference)

The dict layout is the one Trainer.save_model writes (unirec/facility/trainer.py:389-398: config, cur_epoch, cur_step,
best_valid_score, state_dict, optimizer, scheduler), saved with torch.save like accelerator.save does.  Used to show that a
checkpoint produced by the reference loads into the drop-in classes (general.load_model_freely / Trainer.load_model)."""
import os
import sys

import torch

REF = os.environ.get('UNIREC_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden_eval', 'reference_checkpoint_mf.pth')


def main():
    sys.path.insert(0, REF)
    sys.argv = sys.argv[:1]
    from unirec.utils import argument_parser, general
    cfg = argument_parser.parse_arguments(dict(model='MF', dataset='example', exp_name='refckpt', n_users=41, n_items=97, embedding_size=16,
                                               loss_type='bpr', train_file_format='user-item', scheduler='none', has_item_bias=1))
    cfg['device'] = torch.device('cpu')
    general.init_seed(2022)
    model = general.get_class_instance('MF', 'unirec/model')(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    user = torch.randint(1, 41, (8,))
    item = torch.randint(1, 97, (8, 2))
    loss = model(user_id=user, item_id=item, label=None)[0]
    loss.backward()
    opt.step()
    model.eval()
    with torch.no_grad():
        scores = model(user_id=user, item_id=item)[1]
    state = {'config': cfg, 'cur_epoch': 3, 'cur_step': 1, 'best_valid_score': 0.5, 'state_dict': model.state_dict(),
             'optimizer': opt.state_dict(), 'scheduler': None, '_probe': {'user_id': user, 'item_id': item, 'scores': scores}}
    torch.save(state, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
