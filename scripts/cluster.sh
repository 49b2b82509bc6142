#!/bin/bash
# Multi-node GPU cluster helper — the counterpart of the reference's TPU pod tooling (scripts/tpu_commands.sh:184-251:
# list/ips/ssh/copy/setup/launch/check/stop/…), re-thought for B200 nodes reachable over ssh.
#
#   export MIDGPT_HOSTS="node0 node1 node2 node3"     # first host is the rendezvous master
#   source scripts/cluster.sh
#   gpu list | ips | check | df               inspect nodes (nvidia-smi, tmux pane of the run, disk)
#   gpu ssh '<cmd>'                           run a command on every node (fan-out, waits for all)
#   gpu copy                                  rsync this repo to ~/midgpt_b200 on every node
#   gpu setup [dataset_dir]                   scripts/setup_node.sh on every node: verify driver/torch/NVLink, build the sm_100a
#                                             extensions, link the dataset at /mnt/disks/persist, run the smoke (reference: setup.sh)
#   gpu health                                ECC / throttle / NVLink error counters of every GPU on every node
#   gpu reboot                                reboot every node (reference: tpu reboot)
#   gpu launch <config> <rundir> [flags…]     start training in a detached tmux session on every node (torchrun, 8 ranks/node)
#   gpu stop                                  SIGTERM the run (ranks checkpoint at the next step boundary and exit), then remove the tmux session
#   gpu maintain <host>                       fault injection: SIGKILL one rank on <host> (run then resumes from the last checkpoint)
#   gpu retry_launch <config> <rundir>        relaunch until it exits 0 (restart-from-checkpoint recovery loop)
GPU_REPO_DIR=${GPU_REPO_DIR:-midgpt_b200}
GPU_PORT=${GPU_PORT:-29400}
GPU_PER_NODE=${GPU_PER_NODE:-8}

_gpu_hosts() { echo ${MIDGPT_HOSTS:?set MIDGPT_HOSTS to a space-separated host list}; }
_gpu_each() { local cmd="$1"; for h in $(_gpu_hosts); do ssh -o StrictHostKeyChecking=no "$h" "$cmd" 2>&1 | sed "s/^/[$h] /" & done; wait; }

gpu() {
  local sub=$1; shift
  local hosts=($(_gpu_hosts)); local master=${hosts[0]}; local nnodes=${#hosts[@]}
  case $sub in
    list)   _gpu_each "nvidia-smi --query-gpu=index,name,memory.used,utilization.gpu --format=csv,noheader" ;;
    ips)    for h in ${hosts[@]}; do getent hosts $h | awk -v h=$h '{print h, $1}'; done ;;
    df)     _gpu_each "df -h . | tail -1" ;;
    ssh)    _gpu_each "$*" ;;
    copy)   for h in ${hosts[@]}; do rsync -az --delete --exclude .git --exclude outputs --exclude gpurun_out ./ $h:$GPU_REPO_DIR/ & done; wait ;;
    setup)  _gpu_each "cd $GPU_REPO_DIR && bash scripts/setup_node.sh $1" ;;
    health) _gpu_each "nvidia-smi --query-gpu=index,ecc.errors.uncorrected.volatile.total,clocks_event_reasons.hw_slowdown,temperature.gpu,power.draw --format=csv,noheader; nvidia-smi nvlink -e 2>/dev/null | grep -ci error | xargs -I{} echo 'nvlink error lines: {}'" ;;
    reboot) _gpu_each "sudo reboot" ;;
    check)  _gpu_each "tmux capture-pane -pt midgpt 2>/dev/null | tail -n 5 || echo 'no session'" ;;
    launch)
      local config=$1 rundir=$2; shift 2
      local i=0
      for h in ${hosts[@]}; do
        ssh $h "cd $GPU_REPO_DIR && tmux new -d -s midgpt \"python -m torch.distributed.run --nnodes=$nnodes --node-rank=$i --nproc-per-node=$GPU_PER_NODE --master-addr=$master --master-port=$GPU_PORT launch.py --config=$config --rundir=$rundir --multihost $* 2>&1 | tee -a launch_$i.log\"" &
        i=$((i+1))
      done; wait ;;
    stop)   # graceful first: SIGTERM to the run's process group -> every rank checkpoints at the next step boundary and exits 0
            # (Trainer.fit's cooperative preemption), then the session is removed
            _gpu_each "pid=\$(tmux list-panes -t midgpt -F '#{pane_pid}' 2>/dev/null | head -1); [ -n \"\$pid\" ] && kill -TERM -- -\$(ps -o pgid= \$pid | tr -d ' ') 2>/dev/null; for i in \$(seq ${GPU_STOP_GRACE:-60}); do tmux has-session -t midgpt 2>/dev/null || break; sleep 1; done; tmux kill-session -t midgpt 2>/dev/null; true" ;;
    maintain)
      local host=$1
      ssh $host "pid=\$(nvidia-smi --query-compute-apps=pid --format=csv,noheader | head -1); [ -n \"\$pid\" ] && kill -9 \$pid && echo killed \$pid" ;;
    retry_launch)
      local config=$1 rundir=$2; shift 2
      until gpu launch $config $rundir "$@" && gpu wait; do echo "run failed — relaunching in 30 s (resumes from the last checkpoint)"; gpu stop; sleep 30; done ;;
    wait)   while _gpu_each "tmux has-session -t midgpt 2>/dev/null && echo alive" | grep -q alive; do sleep 20; done ;;
    *) echo "usage: gpu {list|ips|df|ssh|copy|setup|health|reboot|check|launch|stop|maintain|retry_launch|wait}" ;;
  esac
}
