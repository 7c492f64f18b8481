"""Thin epoch driver with the reference's entry-point contract (`train.py:20-74`, `xnrs/training.py:118-189`).

    python -m xnrs_b200.train --config config/mind_small_CL.yml

reads the reference's YAML configs UNCHANGED (same keys: data paths, model, hyper-parameters, `n_epochs`, `test_freq`,
`ckpt_freq`, `dir`, `name`, `random_seed`, `debug` behaviour), builds the model with `make_model(cfg)`, and runs

  * training epochs over int32 index batches assembled by `mind_io` from the reference's on-disk formats (the pickled news
    frame of `make_mind_dataset.py` and the behaviours CSV of `mind.py:186-199`) — `ContrastiveRankingTrainer` like the
    reference hard-codes (`train.py:71`), except NPA, which has no contrastive hook in the reference either;
  * the full-catalogue evaluation (`CatalogueEvaluator`) every `test_freq` epochs instead of the reference's
    one-impression-per-step loop — same per-impression metrics, unweighted epoch means (`training.py:257-266`);
  * save-only checkpoints in the reference's format `{config, model_name, state_dict}` at
    `<dir>/<name>/checkpoints/ckpt_<epoch>` (`training.py:73-83`) — the state_dict keys are the reference's, so either side
    can load the other's files.

Host orchestration only: no tensor math happens here.  wandb logging is not reproduced.
"""
from __future__ import annotations

import argparse
import os
import random
from typing import Dict, List, Optional, Sequence

import torch

from . import mind_io
from .data import TitleStore
from .evaluation import CatalogueEvaluator
from .models import make_model
from .training import Cfg, ContrastiveRankingTrainer, MSERankingTrainer


def load_config(path: str, **overrides) -> Cfg:
    """yaml.full_load like the reference (`train.py:23`; a duplicated key keeps its last value, e.g. NAML's random_seed)"""
    import yaml
    cfg = yaml.full_load(open(path, 'r'))
    cfg.update(overrides)
    return Cfg(cfg)


def _raw(cfg: Cfg) -> dict:
    return cfg._c if isinstance(cfg._c, dict) else dict(cfg._c)


class MindData:
    """the reference's (train_ds, test_ds) pair (`make_mind_data`, mind.py:13-80) in index form"""

    def __init__(self, cfg: Cfg, news_feat: Optional[Dict[str, dict]] = None, train_sessions: Optional[List[dict]] = None,
                 test_sessions: Optional[List[dict]] = None, test_news_feat: Optional[Dict[str, dict]] = None):
        text = list(cfg.get('text_features', ['title_emb']))
        catg = list(cfg.get('catg_features', []) or [])
        min_len = int(cfg.get('min_hist_len', 1) or 1)

        def tables(feat_dict, path):
            if feat_dict is not None:
                return mind_io.NewsTables.from_news_dict(feat_dict, text, catg)
            return mind_io.NewsTables.from_pickle(path, text, catg) if path else None

        def sessions(given, path):
            s = given if given is not None else (mind_io.read_behaviors(path) if path else None)
            return None if s is None else [x for x in s if len(x['history']) >= min_len]

        self.text_features, self.catg_features = text, catg
        self.train_tables = tables(news_feat, cfg.get('train_news_data_path'))
        self.test_tables = tables(test_news_feat if test_news_feat is not None else (news_feat if cfg.get('test_news_data_path') is None else None),
                                  cfg.get('test_news_data_path'))
        self.train_sessions = sessions(train_sessions, cfg.get('train_user_data_path'))
        self.test_sessions = sessions(test_sessions, cfg.get('test_user_data_path'))
        all_sessions = (self.train_sessions or []) + (self.test_sessions or [])
        self.themes = mind_io._theme_ids(all_sessions)            # numbered once, globally (ranks / epochs agree)


def index_batch(tables: mind_io.NewsTables, stores: Dict[str, TitleStore], raw: Dict[str, torch.Tensor], device,
                text_features: Sequence[str], catg_features: Sequence[str]) -> dict:
    """reference batch-dict schema carrying IndexedTitles instead of dense (x, m) pairs"""
    hist_ids = raw['hist_ids'].to(device, non_blocking=True)
    cand_ids = raw['cand_ids'].to(device, non_blocking=True)
    hist = {f: stores[f].index(hist_ids) for f in text_features}
    cand = {f: stores[f].index(cand_ids) for f in text_features}
    for f in catg_features:
        col = tables.categorical[f].to(device)
        hist[f], cand[f] = col[hist_ids.long()], col[cand_ids.long()]
    return {'user_features': {'history': hist, 'other': {'user_index': raw['user_index'].to(device, non_blocking=True)}},
            'candidate_features': cand, 'targets': raw['targets'].to(device, non_blocking=True),
            'main_theme': raw['main_theme'].to(device, non_blocking=True)}


def _stores(tables: mind_io.NewsTables, text_features, device) -> Dict[str, TitleStore]:
    table = tables.token_table.to(device)
    return {f: TitleStore(table, tables.tokens[f].to(device)) for f in text_features}


def save_scores(cfg: Cfg, epoch: int, targets, scores, stats: dict) -> str:
    """training.py:85-95: `<dir>/<name>/predictions/predictions_<epoch>` = {targets, scores, stats} (numpy arrays)"""
    path = os.path.join(cfg.dir, cfg.name, 'predictions')
    os.makedirs(path, exist_ok=True)
    fname = os.path.join(path, f'predictions_{epoch}')
    torch.save({'targets': targets, 'scores': scores, 'stats': stats}, fname)
    return fname


def evaluate(model, cfg: Cfg, tables: mind_io.NewsTables, sessions: Sequence[dict], device, debug: bool = False,
             epoch: Optional[int] = None) -> dict:
    """`_test_iteration` + `_after_test_iteration` (training.py:131-142, 194-303) as one full-catalogue pass: the epoch
    means of every `_test_step` metric (ranking + thresholded), the summed confusion matrix, and — when `epoch` is given —
    the reference's per-epoch prediction dump (all scores / targets, per-impression auc / mrr / ndcg)."""
    stores = _stores(tables, list(cfg.get('text_features', ['title_emb'])), device)
    text = list(cfg.get('text_features', ['title_emb']))
    cat = tables.categorical.get('category_index')
    sub = tables.categorical.get('subcategory_index')
    ev = CatalogueEvaluator(model, stores[text[0]], cat, sub, stores.get('abstract_emb'))
    ev.binary_metrics = True
    imp = mind_io.eval_impressions(sessions[:1] if debug else sessions, tables, int(cfg.hist_len))
    model.eval()
    out = ev.evaluate(imp, return_per_impression=epoch is not None)
    res = {k: out[k] for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10', 'acc', 'rec', 'prec', 'conf', 'impressions')}
    if epoch is not None and out.get('per_impression') is not None:
        per = out['per_impression'].cpu().numpy()
        res['predictions'] = save_scores(cfg, epoch, imp['targets'].numpy(), out['scores'].cpu().numpy(),
                                         {'auc': per[:, 0], 'mrr': per[:, 1], 'ndcg@5': per[:, 2], 'ndcg@10': per[:, 3]})
    return res


def save_checkpoint(cfg: Cfg, model, epoch: int) -> str:
    """training.py:73-83"""
    path = os.path.join(cfg.dir, cfg.name, 'checkpoints')
    os.makedirs(path, exist_ok=True)
    fname = os.path.join(path, f'ckpt_{epoch}')
    torch.save({'config': dict(_raw(cfg)), 'model_name': cfg.name, 'state_dict': model.state_dict()}, fname)
    return fname


def load_model_from_ckpt(path: str, device: Optional[str] = None):
    """xnrs/models/utils.py:14-21: rebuild the model from a checkpoint's own config and load its weights"""
    ck = torch.load(path, map_location='cpu', weights_only=False)
    cfg = Cfg(dict(ck['config'], **({'device': device} if device else {})))
    model = make_model(cfg)
    model.load_state_dict(ck['state_dict'], strict=True)
    return model.to(cfg.get('device', 'cuda:0')), cfg


def train(cfg_path: str, debug: bool = False, data: Optional[MindData] = None, **overrides) -> dict:
    """`train(cfg_path, debug)` of the reference (train.py:20-74).  Returns {'train_loss': [...], 'test': [...]} per epoch."""
    cfg = load_config(cfg_path, **overrides)
    _raw(cfg)['debug'] = bool(debug)
    if debug:
        _raw(cfg)['name'] = 'debug_run'
    os.makedirs(os.path.join(cfg.dir, cfg.name), exist_ok=True)
    torch.manual_seed(int(cfg.random_seed))
    random.seed(int(cfg.random_seed))
    if cfg.get('dataset', 'mind') != 'mind':
        raise ValueError('unknown dataset (only the MIND formats are implemented; Adressa is out of scope, SURVEY §2 row 21)')
    model = make_model(cfg)
    data = data if data is not None else MindData(cfg)
    device = torch.device(cfg.get('device', 'cuda:0'))
    trainer_cls = MSERankingTrainer if str(cfg.model).lower() == 'npa' else ContrastiveRankingTrainer
    trainer = trainer_cls(cfg, model)
    text, catg = data.text_features, data.catg_features
    history = {'train_loss': [], 'test': [], 'checkpoints': []}
    n_epochs, B = int(cfg.n_epochs), int(cfg.batch_size)
    rng = random.Random(int(cfg.random_seed))
    stores = _stores(data.train_tables, text, device) if data.train_tables is not None else None

    current = [0]

    def run_test():
        if data.test_tables is not None and data.test_sessions:
            history['test'].append(evaluate(model, cfg, data.test_tables, data.test_sessions, device, debug, epoch=current[0]))
            print('test:', history['test'][-1])

    for epoch in range(n_epochs):
        current[0] = epoch
        print(f'\n Epoch {epoch}:')
        if stores is not None and data.train_sessions:
            model.train()
            order = list(range(len(data.train_sessions)))
            if cfg.get('shuffle_data', True):
                rng.shuffle(order)
            losses = []
            for lo in range(0, len(order) - B + 1, B):                       # drop_last like the reference loader
                raw = mind_io.train_batch(data.train_sessions, data.train_tables, order[lo:lo + B], int(cfg.hist_len),
                                          int(cfg.n_negatives), rng=rng, themes=data.themes)
                out = trainer._train_step(index_batch(data.train_tables, stores, raw, device, text, catg))
                losses.append(out['loss'])
                if debug:
                    print('debugging - interrupting after first step')
                    break
            epoch_loss = float(torch.stack([l.reshape(()) for l in losses]).mean()) if losses else float('nan')
            history['train_loss'].append(epoch_loss)
            print('train loss: ', epoch_loss)
            freq = cfg.get('ckpt_freq', None)
            if freq is not None and (epoch % int(freq) == 0 or epoch == n_epochs - 1):
                history['checkpoints'].append(save_checkpoint(cfg, model, epoch))
        if (epoch + 1) % int(cfg.get('test_freq', 1)) == 0 or epoch == n_epochs - 1:
            run_test()
        if debug:
            print('debugging - interrupting after first epoch')
            break
    if n_epochs == 0:
        run_test()
    history['model'] = model
    return history


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', '-c', default='config/mind_small.yml', help='path to a reference config file')
    ap.add_argument('--debug', action='store_true')
    a = ap.parse_args()
    train(a.config, debug=a.debug)
