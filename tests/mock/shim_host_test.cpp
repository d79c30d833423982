// Host-logic checks of tako_b200::GpuCore against the test double of the C ABI (fake_hqsched.cpp): what the shim
// does around a tick — interning, handle mapping, batched pushes, applying the result to its worker mirror,
// resource return, worker removal (min_utilization is enforced inside the library's tick: tests/test_gpu_edges.py).  Returns the number of failed checks.
#include "../../include/tako_shim.hpp"

#include <cstdio>
#include <stdexcept>

using namespace tako_b200;

static int failed = 0;
static void check(bool ok, const char* what) {
    if (!ok) { ++failed; std::fprintf(stderr, "FAILED: %s\n", what); }
}
static ResourceRequestVariants cpus(uint64_t n) {
    ResourceRequest rq;
    rq.entries.push_back({0, false, n * FRACTIONS_PER_UNIT});
    return ResourceRequestVariants{{rq}};
}

int main() {
    {   // interning + per-worker priority order + free vectors + resource return
        GpuCore core(2, 0);
        const ResourceRqId c1 = core.get_or_create_resource_rq_id(cpus(1)), c2 = core.get_or_create_resource_rq_id(cpus(2));
        check(core.get_or_create_resource_rq_id(cpus(1)) == c1 && c1 != c2, "interning");
        ResourceRequestVariants all_gpu = cpus(1);
        all_gpu.variants[0].entries.push_back({1, true, 0});
        const ResourceRqId ca = core.get_or_create_resource_rq_id(all_gpu);
        core.on_new_worker(7, {4 * FRACTIONS_PER_UNIT, 2 * FRACTIONS_PER_UNIT});
        core.on_new_worker(3, {2 * FRACTIONS_PER_UNIT, 0});
        core.add_ready_task(TaskId{1, 10}, c2, priority_from_user(1));
        core.add_ready_task(TaskId{1, 11}, c1, priority_from_user(5));
        core.add_ready_task(TaskId{1, 12}, ca, priority_from_user(3));
        core.add_ready_task(TaskId{1, 13}, c2, priority_from_user(0));
        check(core.stats().kernel_launches == 0, "pushes are batched until the tick");
        WorkerTaskMapping m = core.run_scheduling();
        check(core.stats().kernel_launches == 4, "one batched push of four tasks");
        // first-fit in worker-id order: worker 3 (2 cpus) gets task 11 (prio 5); task 12 (All gpus) only fits worker 7;
        // task 10 (2 cpus, prio 1) -> worker 7; task 13 (2 cpus) does not fit any more (worker 7 has 1 cpu left, worker 3 has 1)
        check(m.n_assigned() == 3, "three of four tasks placed");
        check(m.workers.count(3) && m.workers[3].assigned.size() == 1 && m.workers[3].assigned[0].first == (TaskId{1, 11}), "worker 3 list");
        check(m.workers.count(7) && m.workers[7].assigned.size() == 2 && m.workers[7].assigned[0].first == (TaskId{1, 12}) &&
              m.workers[7].assigned[1].first == (TaskId{1, 10}), "worker 7 list is priority-descending");
        check(core.free_resources(7)[0] == 1 * FRACTIONS_PER_UNIT && core.free_resources(7)[1] == 0, "All consumed the whole resource");
        core.on_task_finished(TaskId{1, 12});
        check(core.free_resources(7)[0] == 2 * FRACTIONS_PER_UNIT && core.free_resources(7)[1] == 2 * FRACTIONS_PER_UNIT, "finish restores All to the total");
        m = core.run_scheduling();
        check(m.n_assigned() == 1 && m.workers[7].assigned[0].first == (TaskId{1, 13}), "the waiting task runs after resources return");
        check(core.run_scheduling().n_assigned() == 0, "a second tick emits nothing");
    }
    {   // cancel before and after the flush, worker removal requeues its tasks
        GpuCore core(1, 0);
        const ResourceRqId c1 = core.get_or_create_resource_rq_id(cpus(1));
        core.on_new_worker(1, {1 * FRACTIONS_PER_UNIT});
        core.add_ready_task(TaskId{2, 1}, c1, priority_from_user(0));
        core.add_ready_task(TaskId{2, 2}, c1, priority_from_user(0));
        core.remove_ready_task(TaskId{2, 1});                          // still in the host batch
        check(core.stats().n_segments == 0, "cancel of an unflushed task does not reach the library");
        WorkerTaskMapping m = core.run_scheduling();
        check(m.n_assigned() == 1 && m.workers[1].assigned[0].first == (TaskId{2, 2}), "cancelled task is never scheduled");
        core.add_ready_task(TaskId{2, 3}, c1, priority_from_user(0));
        check(core.run_scheduling().n_assigned() == 0, "no capacity left");
        core.remove_ready_task(TaskId{2, 3});
        check(core.stats().n_segments == 1, "cancel of a flushed task goes to hqs_ready_remove");
        core.on_new_worker(2, {1 * FRACTIONS_PER_UNIT});
        core.on_remove_worker(1);                                      // task {2,2} was running there
        m = core.run_scheduling();
        check(m.n_assigned() == 1 && m.workers.count(2) && m.workers[2].assigned[0].first == (TaskId{2, 2}), "tasks of a lost worker are rescheduled");
    }
    {   // error behaviour: invalid requests throw, like the reference's panics
        GpuCore core(1, 0);
        bool threw = false;
        try { ResourceRequest rq; rq.entries.push_back({0, false, 0}); core.get_or_create_resource_rq_id(ResourceRequestVariants{{rq}}); }
        catch (const std::invalid_argument&) { threw = true; }
        check(threw, "zero amount is refused");
        threw = false;
        try { core.add_ready_task(TaskId{9, 9}, 77, 0); } catch (const std::invalid_argument&) { threw = true; }
        check(threw, "unknown request id is refused");
    }
    {   // proactive filling, retract + redirect, on_retract_response, RunningPrefilled (mapping.rs:63-101, 156-230, 255-288;
        // reactor.rs:263-345, 452-498): the shim's bookkeeping of kind 1 / kind 2 records
        GpuCore core(1, 0);
        core.set_scheduler_config(0, 10);
        const ResourceRqId c1 = core.get_or_create_resource_rq_id(cpus(1));
        core.on_new_worker(1, {2 * FRACTIONS_PER_UNIT});
        std::vector<TaskId> ids;
        for (uint32_t k = 1; k <= 4; ++k) ids.push_back(TaskId{3, k});
        core.on_new_tasks(ids);
        for (const TaskId& t : ids) core.add_ready_task(t, c1, priority_from_user(0));
        WorkerTaskMapping m = core.run_scheduling();
        check(m.workers[1].assigned.size() == 2 && m.workers[1].prefills.size() == 2 && m.workers[1].retracts.empty(),
              "two tasks run, the other two are prefilled behind them");
        check(m.workers[1].assigned[0].first == (TaskId{3, 1}) && m.workers[1].prefills[0] == (TaskId{3, 3}), "ascending TaskId inside the class");
        check(core.n_prefilled(1) == 2 && core.free_resources(1)[0] == 0, "prefilled tasks take no resources");
        check(core.run_scheduling().workers.empty(), "nothing changes without an event");
        // a second worker appears: the prefilled tasks are the only ready ones, they move (RetractTasks to worker 1,
        // redirect to worker 2)
        core.on_new_worker(2, {2 * FRACTIONS_PER_UNIT});
        m = core.run_scheduling();
        check(m.workers.count(1) && m.workers[1].retracts.size() == 2 && m.workers[1].assigned.empty(), "RetractTasks goes to the holder");
        check(!m.workers.count(2) || m.workers[2].assigned.empty(), "the new worker gets the tasks only after the retract response");
        check(core.redirects().size() == 2 && core.n_prefilled(1) == 0, "SchedulerState::redirects holds both tasks");
        check(core.free_resources(2)[0] == 0, "the target's resources are taken at once");
        auto to = core.on_retract_response(1, {TaskId{3, 3}, TaskId{3, 4}});
        check(to.size() == 1 && to.count(2) && to[2].size() == 2 && to[2][0].first == (TaskId{3, 3}) && to[2][0].second == 0,
              "the retract response releases the redirected ComputeTasks");
        check(core.redirects().empty(), "redirects are consumed");
        check(core.on_retract_response(1, {TaskId{3, 3}}).empty(), "a second response for the same task is ignored");
        core.on_task_finished(TaskId{3, 3});
        check(core.free_resources(2)[0] == 1 * FRACTIONS_PER_UNIT, "a redirected task returns its resources to the target");
    }
    {   // RunningPrefilled: the worker starts a prefilled task on its own (reactor.rs:263-345)
        GpuCore core(1, 0);
        core.set_scheduler_config(0, 10);
        const ResourceRqId c1 = core.get_or_create_resource_rq_id(cpus(1));
        core.on_new_worker(5, {1 * FRACTIONS_PER_UNIT});
        std::vector<TaskId> ids = {TaskId{4, 1}, TaskId{4, 2}, TaskId{4, 3}};
        core.on_new_tasks(ids);
        for (const TaskId& t : ids) core.add_ready_task(t, c1, priority_from_user(0));
        WorkerTaskMapping m = core.run_scheduling();
        check(m.workers[5].assigned.size() == 1 && m.workers[5].prefills.size() == 2, "one runs, two are prefilled");
        const uint32_t removes_before = core.stats().n_segments;
        core.on_task_finished(TaskId{4, 1});
        core.on_task_running_prefilled(TaskId{4, 2}, 0);
        check(core.n_prefilled(5) == 1 && core.free_resources(5)[0] == 0, "the started task holds the cpu, one prefill is left");
        check(core.stats().n_segments == removes_before + 1, "the started task leaves the device ready set");
        m = core.run_scheduling();
        check(m.n_assigned() == 0, "the worker is full: the remaining prefilled task stays where it is");
        core.on_task_finished(TaskId{4, 2});
        m = core.run_scheduling();
        check(m.workers.count(5) && m.workers[5].retracts.size() == 1 && core.redirects().size() == 1,
              "a prefilled task that is assigned (here to its own holder) goes through retract + redirect like in the reference");
    }
    std::fprintf(stderr, "shim host test: %d failed\n", failed);
    return failed;
}
