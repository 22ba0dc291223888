"""One training iteration in the reference's order (scripts/train.py:440-493):

1. generator step on the clip batch: ``model(..., use_gt=True)`` -> generator loss (GAN + feature
   matching + flow warp) -> ``optimizer_generator`` (everything except ``acts_to_boxes``, :367);
2. discriminator step on the same ``model_out``: hinge loss on fake / real -> ``optimizer_d_img`` (:522-525);
3. graph step on the long (4 x frames_per_action_graph frames) graph batch: ``model(..., graph_only=True)``
   -> masked smooth-L1 on the boxes -> ``optimizer_graph`` (``acts_to_boxes`` only, :365).

Errors propagate (the reference swallows them, train.py:466-468).  With more than one rank every
backward is followed by the all-reduce(avg) of exactly the gradients its optimiser consumes.

NaN guard (train.py:450-453): the reference skips the generator AND the discriminator update of a clip
batch whose ``GAN_Img`` / ``GAN_Feat`` loss is NaN.  Here the decision stays on the device: the flag is
handed to the fused Adam kernels as their ``found_inf`` input (the AMP hook of torch.optim), which makes
the update - parameters, both moments and the step counter - a no-op without a host synchronisation,
so the iteration can still be captured into a CUDA graph.  Non-fused optimisers (CPU tests) check on the host.

Checkpoints use the reference's dictionary (train.py:528-543, restore :27-55): ``model_state``,
``gans_model_state``, ``d_img_state``, ``d_img_optim_state``, ``optim_state_gen``, ``optim_state_graph``,
``vocab``, ``counters``.
"""
import torch

from . import dist as agdist

_WRAPPED = ('acts_to_boxes', 'acts_to_objs', 'layout_to_video')      # DataParallel wrappers of meta_models.py:16-27


def to_reference_keys(state):
    """Our state-dict keys -> the reference's (its three sub-models sit inside DataParallelWithCallback)."""
    out = {}
    for k, v in state.items():
        head, _, rest = k.partition('.')
        out['%s.module.%s' % (head, rest) if head in _WRAPPED and not rest.startswith('module.') else k] = v
    return out


def from_reference_keys(state):
    return {k.replace('.module.', '.', 1) if k.split('.', 1)[0] in _WRAPPED else k: v for k, v in state.items()}


class Trainer:
    def __init__(self, opt, model, discriminator, gans_model, world=1, fused=True):
        self.opt, self.model, self.discriminator, self.gans_model = opt, model, discriminator, gans_model
        kw = dict(lr=opt.learning_rate, betas=(opt.beta1, 0.999))
        self.fused = bool(fused)
        if fused:
            kw.update(fused=True, capturable=True)
        self.graph_params = list(model.acts_to_boxes.parameters())
        ids = {id(p) for p in self.graph_params}
        self.gen_params = [p for p in model.parameters() if id(p) not in ids]
        self.d_params = list(discriminator.img_discriminator.parameters())
        self.optimizer_graph = torch.optim.Adam(self.graph_params, **kw)
        self.optimizer_generator = torch.optim.Adam(self.gen_params, **kw)
        self.optimizer_d_img = discriminator.optimizer_d_img
        self.t, self.epoch = 0, 0
        self.world = world
        self.overlap_graph_step = bool(getattr(opt, 'overlap_graph_step', False))     # measured on B200: no gain (the 227 KB clusters need whole SMs)
        self._side = None
        import inspect
        self._model_takes_flag = 'boxes_pred_grad' in inspect.signature(model.forward).parameters
        self.buckets = None
        if world > 1:
            self.buckets = {k: agdist.GradBuckets(p) for k, p in
                            (('gen', self.gen_params), ('d', self.d_params), ('graph', self.graph_params))}
        self._skip = None                     # 0-d float32 on the device: 1.0 = skip this clip batch's G and D updates
        if self.fused:
            dev = self.gen_params[0].device
            self._skip = torch.zeros((), device=dev, dtype=torch.float32)
            self.optimizer_generator.found_inf = self._skip
            self.optimizer_d_img.found_inf = self._skip
        # packed TF32 weight copies are rebuilt after every optimiser step (not inferred from backward)
        for o in (self.optimizer_generator, self.optimizer_d_img, self.optimizer_graph):
            o.register_step_post_hook(_invalidate_packs_hook)

    def _sync(self, which):
        if self.buckets is not None:
            self.buckets[which].allreduce()

    def _begin(self, which):
        if self.buckets is not None:
            self.buckets[which].begin()

    def _nan_flag(self, G):
        terms = [G[k].detach() for k in ('GAN_Img', 'GAN_Feat') if k in G]
        if not terms:
            return torch.zeros((), device=G['total_loss'].device, dtype=torch.float32)
        flag = torch.isnan(torch.stack([t.reshape(()) for t in terms])).any().float().reshape(())
        if self.world > 1:                    # a NaN on one rank poisons the all-reduced gradients of all
            import torch.distributed as dist
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return flag

    def _graph_forward_backward(self, graph_batch):
        model, gm = self.model, self.gans_model
        boxes_pred = model(graph_batch.get('imgs'), graph_batch['objs'], graph_batch['triplets'], graph_batch['actions'],
                           boxes_gt=graph_batch['boxes'], test_mode=False, graph_only=True)
        GG = gm(graph_batch, boxes_pred, mode='compute_graph_loss')
        self.optimizer_graph.zero_grad(set_to_none=True)
        self._begin('graph')
        GG['total_loss'].backward()
        return GG

    def iteration(self, batch, graph_batch):
        """batch / graph_batch: dicts with imgs, objs, boxes, triplets, actions (graph_batch needs no
        imgs).  Returns (G_losses, D_losses, G_graph_losses) as dicts of 0-d tensors.

        The graph step's forward and backward (a 15-step recurrence on 2 clips: 32 busy SMs, latency-bound) depend
        on nothing the generator and discriminator steps produce, and neither of those writes ``acts_to_boxes``:
        with ``overlap_graph_step`` they run on a side stream next to the generator / discriminator steps and join
        before the graph gradients are all-reduced and ``optimizer_graph`` steps, so the arithmetic and the update
        order (train.py:440-493) are unchanged."""
        model, gm = self.model, self.gans_model
        kw = {'boxes_pred_grad': False} if self._model_takes_flag else {}
        out = model(batch['imgs'], batch['objs'], batch['triplets'], batch['actions'], boxes_gt=batch['boxes'],
                    test_mode=False, use_gt=True, **kw)
        GG = None
        overlap = self.overlap_graph_step and out[0].is_cuda
        if overlap:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=out[0].device)
            self._side.wait_stream(main)          # after the generator forward: it built this step's weight packs
            with torch.cuda.stream(self._side):
                GG = self._graph_forward_backward(graph_batch)
        G = gm(batch, out, mode='compute_generator_loss')
        flag = self._nan_flag(G)
        skip_host = False
        if self.fused:
            self._skip.copy_(flag)
        else:
            skip_host = bool(flag)
        D = None
        if not skip_host:
            self.optimizer_generator.zero_grad(set_to_none=True)
            self._begin('gen')
            G['total_loss'].backward()
            self._sync('gen')
            self.optimizer_generator.step()

            D = gm(batch, out, mode='compute_discriminator_loss')
            self.optimizer_d_img.zero_grad(set_to_none=True)
            self._begin('d')
            D['total_img_loss'].backward()
            self._sync('d')
            self.optimizer_d_img.step()

        if overlap:
            torch.cuda.current_stream().wait_stream(self._side)
        else:
            GG = self._graph_forward_backward(graph_batch)
        self._sync('graph')
        self.optimizer_graph.step()
        self.t += 1
        return G, D, GG

    # ---- checkpoints in the reference's format (train.py:528-543 / :27-55) ---------------------------------
    def state_dict(self):
        gans = {'module.' + k: v for k, v in self.gans_model.state_dict().items()}      # DataParallel wrapper, train.py:375
        return {'model_state': to_reference_keys(self.model.state_dict()),
                'gans_model_state': gans,
                'd_img_state': self.discriminator.img_discriminator.state_dict(),
                'd_img_optim_state': self.optimizer_d_img.state_dict(),
                'optim_state_gen': self.optimizer_generator.state_dict(),
                'optim_state_graph': self.optimizer_graph.state_dict(),
                # the reference builds optimizer_generator from a Python set (train.py:367): its parameter
                # order is arbitrary, so 'optim_state_gen' is only portable between checkpoints written here
                'optim_state_gen_order': 'ag2video_b200.model.parameters() minus acts_to_boxes',
                'vocab': self.opt.vocab, 'counters': {'t': self.t, 'epoch': self.epoch}}

    def load_state_dict(self, ckpt, strict=True):
        """Restore from a checkpoint written by ``state_dict`` or by the reference.  From a reference
        checkpoint everything is consumed except ``optim_state_gen`` (see ``state_dict``), which is
        skipped with the moments left at zero."""
        self.model.load_state_dict(from_reference_keys(ckpt['model_state']), strict=strict)
        self.discriminator.img_discriminator.load_state_dict(ckpt['d_img_state'], strict=strict)
        if 'gans_model_state' in ckpt:
            gans = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in ckpt['gans_model_state'].items()}
            self.gans_model.load_state_dict(gans, strict=False)          # VGG weights of the reference's criterion, if any
        self.optimizer_d_img.load_state_dict(ckpt['d_img_optim_state'])
        if 'optim_state_graph' in ckpt:
            self.optimizer_graph.load_state_dict(ckpt['optim_state_graph'])
        loaded_gen = 'optim_state_gen_order' in ckpt
        if loaded_gen:
            self.optimizer_generator.load_state_dict(ckpt['optim_state_gen'])
        for o in (self.optimizer_generator, self.optimizer_d_img, self.optimizer_graph):
            for group in o.param_groups:                                   # keep this trainer's execution mode
                group['fused'], group['capturable'] = (True, True) if self.fused else (None, False)
                if self.fused:
                    group['foreach'] = None
            if self.fused:                                                 # capturable Adam keeps `step` on the device
                for p, st in o.state.items():
                    if 'step' in st and torch.is_tensor(st['step']):
                        st['step'] = st['step'].to(device=p.device, dtype=torch.float32)
        self.t = int(ckpt.get('counters', {}).get('t', 0))
        self.epoch = int(ckpt.get('counters', {}).get('epoch', 0))
        invalidate_packs()
        return {'optim_state_gen_loaded': loaded_gen}

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)

    def restore_checkpoint(self, path, map_location=None):
        return self.load_state_dict(torch.load(path, map_location=map_location, weights_only=False))


def invalidate_packs():
    from . import spade
    spade.invalidate_packs()


def _invalidate_packs_hook(optimizer, args, kwargs):
    invalidate_packs()
