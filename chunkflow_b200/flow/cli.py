"""``chunkflow``-style command line for the inference hot path.

Chained multi-command group with lazily pulled operator generators, like the reference
(chunkflow/lib/flow.py:44-105).  Only the two commands on the hot path are provided:

    python -m chunkflow_b200.flow.cli create-chunk --size 64 256 256 \
        inference --input-patch-size 20 256 256 --output-patch-overlap 4 64 64 \
                  --num-output-channels 3 --framework b200 --batch-size 12 --mask-output-chunk

``inference`` keeps every flag of the reference operator (chunkflow/flow/flow.py:1853-1893) and adds
the ``b200`` framework choice; ``-f pytorch`` is accepted for the canonical model file.
"""
from functools import update_wrapper
from time import time

import click
import numpy as np

from chunkflow_b200.chunk import Chunk

state = {"dry_run": False, "verbose": 1}


def default_none(ctx, _, value):
    """click turns a missing nargs=3 option into an empty tuple; the operators expect None."""
    return None if value is None or len(value) == 0 else value


@click.group(chain=True)
@click.option("--mip", type=click.INT, default=0, help="default mip level of chunks.")
@click.option("--dry-run/--real-run", default=False, help="dry run or real run.")
@click.option("--verbose/--quiet", default=True, help="print informations or not.")
def main(mip, dry_run, verbose):
    """Compose operators and create your own pipeline (inference hot path only)."""
    state["mip"] = mip
    state["dry_run"] = dry_run
    state["verbose"] = verbose


@main.result_callback()
def process_commands(operators, mip, dry_run, verbose):
    stream = [{"log": {"timer": {}}}]
    for op in operators:
        stream = op(stream)
    tasks = []
    for task in stream:   # pull-driven execution
        tasks.append(task)
    return tasks


def operator(func):
    """Wrap a command so that it returns a closure over the task stream (reference lib/flow.py:82-93)."""
    def new_func(*args, **kwargs):
        def op(stream):
            return func(stream, *args, **kwargs)
        return op
    return update_wrapper(new_func, func)


@main.command("create-chunk")
@click.option("--name", type=str, default="create-chunk", help="name of operator")
@click.option("--size", "-s", type=click.INT, nargs=3, default=(64, 64, 64), help="the size of created chunk")
@click.option("--dtype", type=click.Choice(["uint8", "float32"]), default="uint8", help="the data type of chunk")
@click.option("--pattern", "-p", type=click.Choice(["sin", "random", "zero"]), default="sin")
@click.option("--voxel-offset", "-t", type=click.INT, nargs=3, default=(0, 0, 0), help="offset in voxel number.")
@click.option("--voxel-size", "-e", type=click.INT, nargs=3, default=(1, 1, 1), help="voxel size in nm")
@click.option("--output-chunk-name", "-o", type=str, default="chunk", help="name of created chunk")
@operator
def create_chunk(tasks, name, size, dtype, pattern, voxel_offset, voxel_size, output_chunk_name):
    """Create a fake chunk for easy test (reference flow/flow.py:652-678)."""
    for task in tasks:
        task[output_chunk_name] = Chunk.create(size=size, dtype=np.dtype(dtype), pattern=pattern,
                                               voxel_offset=voxel_offset, voxel_size=voxel_size)
        yield task


@main.command("inference")
@click.option("--name", type=str, default="inference", help="name of this operator")
@click.option("--convnet-model", "-m", type=str, default=None, help="convnet model path or type.")
@click.option("--convnet-weight-path", "-w", type=str, default=None, help="convnet weight path")
@click.option("--input-patch-size", "-s", type=click.INT, nargs=3, required=True, help="input patch size")
@click.option("--output-patch-size", "-z", type=click.INT, nargs=3, default=None, callback=default_none,
              help="output patch size")
@click.option("--output-patch-overlap", "-v", type=click.INT, nargs=3, default=(4, 64, 64), help="patch overlap")
@click.option("--output-crop-margin", type=click.INT, nargs=3, default=None, callback=default_none,
              help="margin size of output chunk cropping.")
@click.option("--patch-num", "-n", default=None, callback=default_none, type=click.INT, nargs=3,
              help="patch number in z,y,x.")
@click.option("--num-input-channels", type=click.INT, default=1, help="number of input channels")
@click.option("--num-output-channels", "-c", type=click.INT, default=3, help="number of output channels")
@click.option("--dtype", "-d", type=click.Choice(["float32", "float16"]), default="float32",
              help="float32: fp16 hi/lo split tensor-core arithmetic with fp32 accumulation (1e-3 parity mode); "
                   "float16: single-pass fp16 tensor cores. The result is float32 either way.")
@click.option("--framework", "-f", type=click.Choice(["universal", "identity", "pytorch", "b200"]), default="universal",
              help="inference framework")
@click.option("--batch-size", "-b", type=click.INT, default=1, help="mini batch size of input patch.")
@click.option("--bump", type=click.Choice(["wu", "zung"]), default="wu", help="bump function type (only support wu now!).")
@click.option("--mask-output-chunk/--no-mask-output-chunk", default=False,
              help="mask output chunk will make the whole chunk like one output patch. "
                   "This will also work with non-aligned chunk size.")
@click.option("--mask-myelin-threshold", "-y", default=None, type=click.FLOAT,
              help="mask myelin if netoutput have myelin channel.")
@click.option("--augment/--no-augment", default=False,
              help="transform the input patch and transform back the output patch")
@click.option("--input-chunk-name", "-i", type=str, default="chunk", help="input chunk name")
@click.option("--output-chunk-name", "-o", type=str, default="chunk", help="output chunk name")
@operator
def inference(tasks, name, convnet_model, convnet_weight_path, input_patch_size, output_patch_size,
              output_patch_overlap, output_crop_margin, patch_num, num_input_channels, num_output_channels, dtype,
              framework, batch_size, bump, mask_output_chunk, mask_myelin_threshold, augment, input_chunk_name,
              output_chunk_name):
    """Perform convolutional network inference for chunks (reference flow/flow.py:1894-1933)."""
    from chunkflow_b200.flow.divid_conquer.inferencer import Inferencer
    with Inferencer(
            convnet_model, convnet_weight_path,
            input_patch_size=input_patch_size, output_patch_size=output_patch_size,
            num_input_channels=num_input_channels, num_output_channels=num_output_channels,
            output_patch_overlap=output_patch_overlap, output_crop_margin=output_crop_margin,
            patch_num=patch_num, framework=framework, dtype=dtype, batch_size=batch_size, bump=bump,
            augment=augment, mask_output_chunk=mask_output_chunk, mask_myelin_threshold=mask_myelin_threshold,
            dry_run=state["dry_run"]) as inferencer:
        for task in tasks:
            if task is not None:
                if "log" not in task:
                    task["log"] = {"timer": {}}
                start = time()
                task[output_chunk_name] = inferencer(task[input_chunk_name])
                task["log"]["timer"][name] = time() - start
                task["log"]["compute_device"] = inferencer.compute_device
                if state["verbose"]:
                    out = task[output_chunk_name]
                    print(f"{name}: {out.shape} in {task['log']['timer'][name]:.3f} s on "
                          f"{task['log']['compute_device']} ({np.prod(out.shape[-3:]) / task['log']['timer'][name] / 1e6:.1f} Mvoxels/s)")
            yield task


if __name__ == "__main__":
    main()
