# CUDA 12.9 toolchain + PyTorch; builds the sm_100a extension in-tree and runs the CPU test-suite.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
RUN apt-get update && apt-get install -y python3 python3-pip python3-venv git && rm -rf /var/lib/apt/lists/*
RUN pip install --break-system-packages torch dill cloudpickle numpy pytest requests
WORKDIR /opt/sparktorch-b200
COPY . .
RUN python3 -c "import __graft_entry__ as g; g.build()"
CMD ["python3", "-m", "pytest", "tests/", "-x", "-q", "-m", "not gpu"]
