"""Synchronous data-parallel training (``mode='synchronous'``, alias ``'async'``).

Public surface mirrors the reference module (sparktorch/distributed.py:53-261):
``mapPartitionsWithIndex``, ``handle_model``, ``train_distributed``.  The body of
``handle_model`` is new: rendezvous → flat-arena engine
(:mod:`sparktorch_b200.runtime.engine`) → rank 0 returns ``(state_dict,
history)``.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Union

from sparktorch_b200._spark import ensure_pyspark

ensure_pyspark()

from pyspark.rdd import RDD, PipelinedRDD  # noqa: E402
from pyspark.taskcontext import BarrierTaskContext  # noqa: E402

from sparktorch_b200.parallel import rendezvous  # noqa: E402
from sparktorch_b200.runtime.engine import TrainConfig, resolve_device, train_partition  # noqa: E402
from sparktorch_b200.util import DataObj, load_base_torch, load_torch_model, rows_to_block  # noqa: E402

SYNC_MODES = ("synchronous", "async", "sync")


def mapPartitionsWithIndex(rdd: RDD, f: Callable, preservesPartitioning: bool = False) -> PipelinedRDD:
    """Barrier-stage ``mapPartitionsWithIndex`` (all tasks start together or not
    at all) – reference distributed.py:53-63."""
    return PipelinedRDD(rdd, f, preservesPartitioning, isFromBarrier=True)


def handle_model(
    index: int,
    data: List[DataObj],
    torch_obj: Union[str, List],
    iters: int = 1000,
    early_stop_patience: int = -1,
    verbose: int = 1,
    mini_batch: int = -1,
    validation_pct: float = 0,
    device: str = "cpu",
    compile_mode: str = None,
    precision: str = "auto",
    engine: str = "auto",
    comm: str = "auto",
    seed: int = None,
    streaming: bool = False,
    time_window=None,
    sampling: str = "random",
    checkpoint_dir: str = None,
    checkpoint_every: int = 0,
    resume: bool = False,
) -> List[Dict]:
    """Body of one barrier task = one replica (reference distributed.py:66-205).

    ``compile_mode`` is accepted for API compatibility and ignored: the
    reference's handling of it is dead code and this framework
    does not use ``torch.compile`` – hot ops are hand-written kernels."""
    context = BarrierTaskContext.get()
    dev = resolve_device(device, context.partitionId())
    if dev.type == "cuda":
        import torch

        torch.cuda.set_device(dev)
    rank, world = rendezvous.init_from_barrier(context, dev)
    try:
        raw = load_torch_model(torch_obj, build_optimizer=False)
        X, Y = rows_to_block(data)
        if X is None:
            raise RuntimeError(f"partition {index} is empty – barrier training needs data in every partition")
        cfg = TrainConfig(
            iters=iters, mini_batch=mini_batch, validation_pct=validation_pct,
            early_stop_patience=early_stop_patience, verbose=verbose, device=str(dev),
            precision=precision, engine=engine, comm=comm, seed=seed, streaming=streaming,
            time_window=tuple(time_window) if time_window else None, sampling=sampling,
            checkpoint_dir=checkpoint_dir, checkpoint_every=checkpoint_every, resume=resume,
        )
        res = train_partition(rank, world, raw.model, raw.criterion, raw.optimizer, raw.optimizer_params, X, Y, cfg, dev)
    finally:
        rendezvous.shutdown()
    if rank == 0:
        return [{"state_dict": res.state_dict, "history": res.history}]
    return []


def train_distributed(
    rdd: RDD,
    torch_obj: str,
    iters: int = 10,
    partition_shuffles: int = 1,
    verbose: int = 1,
    mini_batch: int = -1,
    validation_pct: float = 0.0,
    device: str = "cpu",
    early_stop_patience: int = -1,
    compile_mode: str = None,
    precision: str = "auto",
    engine: str = "auto",
    comm: str = "auto",
    seed: int = None,
    streaming: bool = False,
    return_history: bool = False,
    time_window=None,
    sampling: str = "random",
    checkpoint_dir: str = None,
    checkpoint_every: int = 0,
    resume: bool = False,
    max_stage_retries: int = 1,
):
    """Driver entry: run the barrier stage ``partition_shuffles`` times,
    reshuffling partitions in between and carrying the weights over (the
    reference re-ships the ORIGINAL blob each round, distributed.py:239-259)."""
    torch_loaded, _shapes = load_base_torch(torch_obj)
    state_dict, history = None, None
    for r in range(partition_shuffles):
        blob = torch_loaded

        def make_task(resume_now):
            def task(i, x, blob=blob):
                return handle_model(
                    i, x, torch_obj=blob, iters=iters, verbose=verbose, mini_batch=mini_batch,
                    validation_pct=validation_pct, device=device,
                    early_stop_patience=int(early_stop_patience), compile_mode=compile_mode,
                    precision=precision, engine=engine, comm=comm, seed=seed, streaming=streaming,
                    time_window=time_window, sampling=sampling, checkpoint_dir=checkpoint_dir,
                    checkpoint_every=checkpoint_every, resume=resume_now,
                )

            task._accepts_blocks = True
            return task

        # Barrier-stage semantics: one failed task fails the whole stage; the stage is retried as a
        # whole (like Spark does) and, when checkpointing is on, resumes from the last checkpoint.
        out = None
        for attempt in range(max_stage_retries + 1):
            try:
                out = mapPartitionsWithIndex(rdd, make_task(resume or attempt > 0)).collect()
                break
            except Exception as e:      # sparkshim raises RuntimeError, real PySpark Py4JJavaError / PythonException / SparkException
                if isinstance(e, (KeyboardInterrupt, SystemExit)) or attempt == max_stage_retries or not checkpoint_dir:
                    raise
        state_dict, history = out[0]["state_dict"], out[0]["history"]
        if partition_shuffles - r > 1:
            rdd = rdd.repartition(rdd.getNumPartitions())
            torch_loaded = _reseed_blob(torch_loaded, state_dict)
    return (state_dict, history) if return_history else state_dict


def _reseed_blob(blob: str, state_dict) -> str:
    """New envelope whose model starts from ``state_dict`` (next shuffle round)."""
    from sparktorch_b200.util import TorchObj, torch_decoder, torch_encoder

    raw = torch_decoder(blob)
    model = raw.model(**raw.model_parameters) if raw.is_lazy and raw.model_parameters else (raw.model() if raw.is_lazy else raw.model)
    model.load_state_dict(state_dict)
    crit = raw.criterion() if raw.is_lazy else raw.criterion
    return torch_encoder(TorchObj(model, crit, raw.optimizer, raw.optimizer_params, False, None))
