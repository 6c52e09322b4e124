"""Trainer: the meta-iteration driver.

The reference's own driver (meta_policy_search/meta_trainer.py:7-164) runs UNCHANGED on top of these
classes (put promp_b200/tf_shim on sys.path for its `import tensorflow`).  This
module is an independent driver with the same constructor, loop structure and logged keys for use
where the reference tree is absent (the GPU box, bench.py): sample -> process -> adapt, repeated
num_inner_grad_steps times, one more sample/process, then the outer optimisation.
"""
import time

import numpy as np

from promp_b200.utils import logger


class Trainer(object):
    def __init__(self, algo, env, sampler, sample_processor, policy, n_itr, start_itr=0, num_inner_grad_steps=1,
                 sess=None):
        self.algo, self.env, self.sampler, self.sample_processor = algo, env, sampler, sample_processor
        self.baseline = sample_processor.baseline
        self.policy = policy
        self.n_itr, self.start_itr = n_itr, start_itr
        self.num_inner_grad_steps = num_inner_grad_steps
        self.sess = sess

    def train_iteration(self, itr, log=True):
        t_itr = time.time()
        self.sampler.update_tasks()
        self.policy.switch_to_pre_update()
        all_samples_data = []
        t_sampling = t_proc = t_inner = 0.0
        for step in range(self.num_inner_grad_steps + 1):
            prefix = 'Step_%d-' % step
            t = time.time()
            paths = self.sampler.obtain_samples(log=log, log_prefix=prefix)
            t_sampling += time.time() - t
            t = time.time()
            samples_data = self.sample_processor.process_samples(paths, log='all' if log else False, log_prefix=prefix)
            all_samples_data.append(samples_data)
            t_proc += time.time() - t
            if log:
                self.log_diagnostics(sum(list(paths.values()), []), prefix=prefix)
            t = time.time()
            if step < self.num_inner_grad_steps:
                self.algo._adapt(samples_data)
            t_inner += time.time() - t
        t_outer = time.time()
        self.algo.optimize_policy(all_samples_data, log=log)
        if log:
            logger.logkv('Itr', itr)
            logger.logkv('n_timesteps', self.sampler.total_timesteps_sampled)
            logger.logkv('Time-OuterStep', time.time() - t_outer)
            logger.logkv('Time-InnerStep', t_inner)
            logger.logkv('Time-SampleProc', t_proc)
            logger.logkv('Time-Sampling', t_sampling)
            logger.logkv('ItrTime', time.time() - t_itr)
        return all_samples_data

    # ------------------------------------------------------------------ CUDA-graph replay of the device part
    def capture_graph(self, warmup=3):
        """Capture everything of a meta-iteration that runs on the device (2x rollout + processing, inner adapt,
        K Adam epochs + stats pass: ~40 kernel launches) into ONE CUDA graph.  Returns step(): draws the tasks
        on the host (numpy RNG, as the reference), uploads them, replays the graph.  Requires in-kernel reset
        states (reset_mode='device'), no host logging and a fixed KL coefficient; world_size 1."""
        import torch
        assert self.sampler.reset_mode == 'device', "graph replay needs reset_mode='device'"
        assert not getattr(self.algo, 'adaptive_inner_kl_penalty', False), "adaptive KL coefficient is a host decision"
        self.sampler.enable_device_phase_counter()

        def device_part():
            self.policy.switch_to_pre_update()
            all_samples = []
            for step in range(self.num_inner_grad_steps + 1):
                paths = self.sampler.obtain_samples(log=False)
                samples = self.sample_processor.process_samples(paths, log=False)
                all_samples.append(samples)
                if step < self.num_inner_grad_steps:
                    self.algo._adapt(samples)
            self.algo.optimize_policy(all_samples, log=False)
            return all_samples

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.sampler.update_tasks()
                device_part()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self.sampler.update_tasks()
        with torch.cuda.graph(graph):
            self._graph_samples = device_part()
        self._graph = graph

        def step():
            self.sampler.update_tasks()
            graph.replay()
            return self._graph_samples
        return step

    def train(self):
        start = time.time()
        for itr in range(self.start_itr, self.n_itr):
            logger.log("\n ---------------- Iteration %d ----------------" % itr)
            self.train_iteration(itr)
            logger.logkv('Time', time.time() - start)
            logger.save_itr_params(itr, self.get_itr_snapshot(itr))
            logger.dumpkvs()
        logger.log("Training finished")

    def get_itr_snapshot(self, itr):
        return dict(itr=itr, policy=self.policy, env=self.env, baseline=self.baseline)

    def log_diagnostics(self, paths, prefix):
        self.env.log_diagnostics(paths, prefix)
        self.policy.log_diagnostics(paths, prefix)
        self.baseline.log_diagnostics(paths, prefix)
