"""One training iteration in the reference's order (scripts/train.py:440-493):

1. generator step on the clip batch: ``model(..., use_gt=True)`` -> generator loss (GAN + feature
   matching + flow warp) -> ``optimizer_generator`` (everything except ``acts_to_boxes``, :367);
2. discriminator step on the same ``model_out``: hinge loss on fake / real -> ``optimizer_d_img`` (:522-525);
3. graph step on the long (4 x frames_per_action_graph frames) graph batch: ``model(..., graph_only=True)``
   -> masked smooth-L1 on the boxes -> ``optimizer_graph`` (``acts_to_boxes`` only, :365).

Errors propagate (the reference swallows them, train.py:466-468).  With more than one rank every
backward is followed by the all-reduce(avg) of exactly the gradients its optimiser consumes.
"""
import torch

from . import dist as agdist


class Trainer:
    def __init__(self, opt, model, discriminator, gans_model, world=1, fused=True):
        self.opt, self.model, self.discriminator, self.gans_model = opt, model, discriminator, gans_model
        kw = dict(lr=opt.learning_rate, betas=(opt.beta1, 0.999))
        if fused:
            kw.update(fused=True, capturable=True)
        self.graph_params = list(model.acts_to_boxes.parameters())
        ids = {id(p) for p in self.graph_params}
        self.gen_params = [p for p in model.parameters() if id(p) not in ids]
        self.d_params = list(discriminator.img_discriminator.parameters())
        self.optimizer_graph = torch.optim.Adam(self.graph_params, **kw)
        self.optimizer_generator = torch.optim.Adam(self.gen_params, **kw)
        self.optimizer_d_img = discriminator.optimizer_d_img
        self.buckets = None
        if world > 1:
            self.buckets = {k: agdist.GradBuckets(p) for k, p in
                            (('gen', self.gen_params), ('d', self.d_params), ('graph', self.graph_params))}

    def _sync(self, which):
        if self.buckets is not None:
            self.buckets[which].allreduce()

    def iteration(self, batch, graph_batch):
        """batch / graph_batch: dicts with imgs, objs, boxes, triplets, actions (graph_batch needs no
        imgs).  Returns (G_losses, D_losses, G_graph_losses) as dicts of 0-d tensors."""
        model, gm = self.model, self.gans_model
        out = model(batch['imgs'], batch['objs'], batch['triplets'], batch['actions'], boxes_gt=batch['boxes'],
                    test_mode=False, use_gt=True)
        G = gm(batch, out, mode='compute_generator_loss')
        self.optimizer_generator.zero_grad(set_to_none=True)
        G['total_loss'].backward()
        self._sync('gen')
        self.optimizer_generator.step()

        D = gm(batch, out, mode='compute_discriminator_loss')
        self.optimizer_d_img.zero_grad(set_to_none=True)
        D['total_img_loss'].backward()
        self._sync('d')
        self.optimizer_d_img.step()

        boxes_pred = model(graph_batch.get('imgs'), graph_batch['objs'], graph_batch['triplets'], graph_batch['actions'],
                           boxes_gt=graph_batch['boxes'], test_mode=False, graph_only=True)
        GG = gm(graph_batch, boxes_pred, mode='compute_graph_loss')
        self.optimizer_graph.zero_grad(set_to_none=True)
        GG['total_loss'].backward()
        self._sync('graph')
        self.optimizer_graph.step()
        return G, D, GG
