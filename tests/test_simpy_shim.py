"""Known-answer tests of the oracle kernel (simpy 4.1.1 semantics).

The reference's own unit tests are the conformance suite in the build container
(oracle/run_reference_tests.sh); these few restate the same facts so the kernel
stays pinned on boxes where /root/reference does not exist.
"""

from __future__ import annotations

import pytest
import simpy


def test_store_is_fifo_and_run_until_event_returns_value():
    # reference tests/unit/runtime/test_simulation_runner.py:183-195
    env = simpy.Environment()
    box = simpy.Store(env)
    env.run(until=box.put("first"))
    env.run(until=box.put("second"))
    assert env.run(until=box.get()) == "first"
    assert env.run(until=box.get()) == "second"


def test_run_until_time_is_urgent_and_exclusive():
    env = simpy.Environment()
    seen = []

    def p():
        yield env.timeout(1.0)
        seen.append(env.now)

    env.process(p())
    env.run(until=1.0)          # NORMAL events at exactly `until` are not processed
    assert seen == [] and env.now == 1.0
    env.run(until=1.5)
    assert seen == [1.0]
    with pytest.raises(ValueError):
        env.run(until=1.5)


def test_equal_time_events_pop_in_insertion_order_and_urgent_first():
    env = simpy.Environment()
    order = []

    def a():
        yield env.timeout(1)
        order.append("a")
        env.process(c())        # Initialize is URGENT: runs before b's NORMAL timeout at t=1
        order.append("a2")

    def b():
        yield env.timeout(1)
        order.append("b")

    def c():
        order.append("c")
        yield env.timeout(0)

    env.process(a())
    env.process(b())
    env.run()
    assert order == ["a", "a2", "c", "b"]


def test_container_is_fifo_with_head_of_line_blocking():
    env = simpy.Environment()
    ram = simpy.Container(env, capacity=10, init=10)
    got = []

    def user(name, amount, hold):
        yield ram.get(amount)
        got.append((name, env.now))
        yield env.timeout(hold)
        yield ram.put(amount)

    env.process(user("big0", 8, 5))
    env.process(user("big1", 8, 1))     # blocks until t=5
    env.process(user("small", 1, 1))    # would fit at t=0 but sits behind big1
    env.run()
    assert got == [("big0", 0), ("big1", 5), ("small", 5)]
    assert ram.level == 10


def test_container_get_is_triggered_immediately_but_resumes_later():
    # the fact ServerRuntime relies on (reference runtime/actors/server.py:212-220)
    env = simpy.Environment()
    cpu = simpy.Container(env, capacity=1, init=1)
    log = []

    def first():
        req = cpu.get(1)
        log.append(("first", req.triggered, cpu.level))
        yield req
        log.append(("first-resumed", env.now))

    def second():
        req = cpu.get(1)
        log.append(("second", req.triggered, cpu.level))
        yield env.timeout(0)

    env.process(first())
    env.process(second())
    env.run()
    assert log[:2] == [("first", True, 0), ("second", False, 0)]


def test_step_and_peek():
    env = simpy.Environment()
    env.timeout(2.5)
    assert env.peek() == 2.5
    env.step()
    assert env.now == 2.5 and env.peek() == float("inf")
    with pytest.raises(simpy.EmptySchedule):
        env.step()
