"""GPU memory query (drop-in for reference utils/gpu_tools.py:4-18)."""
import subprocess


def get_gpu_memory(gpu_id):
    """memory in use on GPU `gpu_id`, MiB, as nvidia-smi reports it."""
    result = subprocess.check_output(['nvidia-smi', '--query-gpu=memory.used', '--format=csv,nounits,noheader'])
    gpu_memory = [int(x) for x in result.decode().strip().split('\n')]
    return dict(zip(range(len(gpu_memory)), gpu_memory))[gpu_id]
