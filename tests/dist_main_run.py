"""Worker: `main.run` end to end under torch.distributed.run (one process per GPU, row-sharded tables set up by main.run itself):
device batch builder with per-rank sample streams of equal length -> SASRec with dropout -> row-sharded training step (CUDA graph) ->
one-vs-all evaluation over the sharded table -> best checkpoint (full tables re-assembled) -> test evaluation from the checkpoint."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main(root, out):
    sys.argv = sys.argv[:1]
    from unirec_b200.main import main as entry
    res = entry.run(dict(model='SASRec', dataset='example', dataloader='SeqRecDataset', exp_name='e2e_dist', dataset_path=root,
                         output_path=out, n_users=60, n_items=200, embedding_size=32, hidden_size=32, n_layers=1,
                         n_heads=2, inner_size=64, max_seq_len=8, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
                         loss_type='softmax', n_sample_neg_train=20, train_file_format='user-item', valid_file_format='user-item',
                         test_file_format='user-item', user_history_filename='user_history', user_history_file_format='user-item',
                         history_mask_mode='autoregressive', seq_last=1, epochs=30, batch_size=32, learning_rate=0.01,
                         scheduler='none', early_stop=0, num_workers=0, valid_protocol='one_vs_all', test_protocol='one_vs_all',
                         metrics="['hit@10', 'ndcg@10', 'group_auc']", key_metric='group_auc', gpu_id=-1))
    if int(os.environ.get('RANK', '0')) == 0:
        print('MAIN_RUN_RESULT ' + json.dumps({k: float(v) for k, v in res.items()}))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
