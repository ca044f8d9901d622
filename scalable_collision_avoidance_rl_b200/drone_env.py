"""Drop-in replacement for the reference's ``drone_env`` module.

Put this repository's root on ``sys.path`` ahead of the reference and the
reference's own drivers (``train_problem.py:4-5``, ``benchmark_agent.py:5-6``)
import this module unchanged: same class ``drones``, same constructor, same
attributes, same ``step()`` 6-tuple, same module-level names.

What differs is where the work happens: ``drones.step()`` / ``rewards()`` run the
fused CUDA kernel of ``libdronestep.so`` (float64 instantiation, E = 1) through
the C ABI, instead of the reference's NumPy loops (reference drone_env.py:214-401).
Host-side setup (formation, d_safety, delta clip, lattice start; reference
drone_env.py:55-212) stays in host float64 and consumes Python's ``random`` /
``numpy.random`` streams exactly as the reference does.

There is no CPU fallback: without a CUDA device the constructor raises.
"""
from __future__ import annotations

import numpy as np

from . import formation
from .batched import BatchedDrones
from .formation import num_to_rgb  # noqa: F401  (re-exported, train_problem.py:150)

# module constants, read at call time like the reference's globals (drone_env.py:26-30)
dim = 2
dt = 0.05
max_time_steps = 200

LIGHT_RED, LIGHT_GREEN, BLACK, WHITE = '#FFC4CC', '#95FD99', '#000000', '#FFFFFF'
LIGHT_PURPLE, LIGHT_ORANGE, YELLOW, BLUE = '#E8D0FF', '#FAE0C3', "#FFFF00", '#98F5FF'


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception as exc:  # plotting is off the hot path; fail only when actually used
        raise RuntimeError("matplotlib is required for the plotting helpers of drone_env") from exc


class drones:
    """Formation-control environment, API of the reference class (drone_env.py:53)."""

    def __init__(self, n_agents: int, n_obstacles: int, grid: list, end_formation: str, k_closest=2,
                 deltas: np.ndarray = None, simplify_zstate=False) -> None:
        self.n_agents = n_agents
        self.grid = grid
        self.goal = self.grid
        self.k_closest = k_closest
        self.simplify_zstate = simplify_zstate
        self.internal_t = 0
        self.collision_weight = 0.2
        self.drone_radius = np.ones(n_agents) * formation.DRONE_RADIUS
        self.A = np.eye(dim)
        self.B = np.eye(dim) * dt
        self.obstacles = self.create_obstacles(n_obstacles)
        self.end_points, self.d_safety = self.generate_formation(end_formation)
        self.deltas, clipped = formation.clip_deltas(deltas, self.d_safety)
        if clipped:
            print("Some deltas are greater than the final minimum distance between end positions. "
                  "Using minimum distance between end positions for those cases instead.",
                  f"deltas = {self.deltas}")
        self._engines = {}
        self.state, self.z_states = self.init_agents(n_agents)

    # ------------------------------------------------------------------ setup (host)
    def reset(self, renew_obstacles=True):
        self.state, self.z_states = self.init_agents(self.n_agents)
        self.internal_t = 0
        if renew_obstacles == True:  # noqa: E712  (reference semantics, drone_env.py:101)
            self.obstacles = self.create_obstacles(self.n_obstacles)

    def __str__(self):
        print("Grid size: [x_lim, y_lim]\n", self.grid)
        print("State: [x, y, vx, vy, r]\n", self.state)
        print(f"z_sattes for k_closest = {self.k_closest}: simplify? {self.simplify_zstate}")
        print("safety distance for each agent:\n", self.d_safety)
        print("Deltas disk radius for each agent: \n", self.deltas)
        print(f"Collision cost weight (per unit of time) = {self.collision_weight} ")
        return ""

    def generate_formation(self, end_formation):
        pts = formation.end_formation(end_formation, self.n_agents, self.grid)
        return pts, formation.safety_distances(pts, self.drone_radius)

    def create_obstacles(self, n_obstacles):
        # consumes numpy's global stream like the reference (drone_env.py:155-169)
        self.n_obstacles = n_obstacles
        hi = 0.1 * np.max(self.grid)
        lo = 0.05 * hi
        obstacles = np.random.rand(n_obstacles, dim + 1)
        obstacles[:, 0] = obstacles[:, 0] * self.grid[0]
        obstacles[:, 1] = obstacles[:, 1] * self.grid[1]
        obstacles[:, dim] = obstacles[:, dim] * (hi - lo) + lo
        return obstacles

    def init_agents(self, n_agents):
        self.n_agents = n_agents
        self.global_state_space = n_agents * (2 * dim + 1)
        self.local_state_space = formation.local_state_space(self.k_closest, self.simplify_zstate)
        self.global_action_space = n_agents * dim
        self.local_action_space = dim
        state = np.zeros([n_agents, 5])
        state[:, 4] = formation.DRONE_RADIUS
        state[:, 0:dim] = formation.sample_start_reference_stream(n_agents, self.grid)
        _, _, z_states, Ni, _ = self.rewards(state, self.end_points, self.n_agents, self.d_safety, self.deltas)
        self.Ni = Ni
        return state, z_states

    # ------------------------------------------------------------------ device engine
    def _engine(self, end_points, d_safety, deltas, radii=None) -> BatchedDrones:
        end_points = np.asarray(end_points, np.float64)
        d_safety = np.asarray(d_safety, np.float64)
        deltas = np.asarray(deltas, np.float64)
        radii = None if radii is None else np.ascontiguousarray(radii, np.float64)
        key = (end_points.tobytes(), d_safety.tobytes(), deltas.tobytes(), self.k_closest,
               bool(self.simplify_zstate), d_safety.shape[0], None if radii is None else radii.tobytes())
        eng = self._engines.get(key)
        if eng is None:
            n = d_safety.shape[0]
            eng = BatchedDrones(1, n, self.grid, "O", self.k_closest, None, self.simplify_zstate,
                                constants=(end_points, d_safety, deltas, radii),
                                start_positions=np.zeros((1, n, 2)), warn=False)
            if len(self._engines) > 8:
                self._engines.clear()
            self._engines[key] = eng
        eng.collision_weight = self.collision_weight   # attribute may be mutated (train_problem.py:31)
        eng.dt, eng.max_time_steps = dt, max_time_steps
        return eng

    def _unpack_obs(self, z, Ni):
        n = z.shape[0]
        z_states = [np.array(z[i]) for i in range(n)]
        Ni_list = [[int(j) for j in Ni[i] if j >= 0] for i in range(n)]
        return z_states, Ni_list

    def rewards(self, state, end_points, n_agents, d_safety, deltas):
        """Fused distance/reward/observation evaluation of `state` (drone_env.py:260-293)."""
        eng = self._engine(end_points, d_safety, deltas)
        self._synced = None                              # the device state is overwritten below
        eng.set_state(np.asarray(state, np.float64)[None], internal_t=self.internal_t)
        out = eng.observe_host()                          # one transfer of the whole result block, one synchronise
        z_states, Ni = self._unpack_obs(out["z"][0], out["Ni"][0])
        return np.array(out["r"][0]), np.int64(out["nc"][0]), z_states, Ni, np.array(out["tr"][0])

    # ------------------------------------------------------------------ hot path
    def step(self, actions):
        """One environment step (drone_env.py:214-258): returns
        (state [aliased], z_states, r_vec, n_collisions, finished, true_r_vec)."""
        n = self.n_agents
        act = np.empty((n, dim))
        for i in range(n):
            act[i] = np.asarray(actions[i], np.float64).reshape(-1)[:dim]
        eng = self._engine(self.end_points, self.d_safety, self.deltas)
        # the host array is the source of truth (callers may have edited it in place): upload it unless
        # it still is, bit for bit, what the device left after the previous step
        sync = getattr(self, "_synced", None)
        if not (sync is not None and sync[0] is eng and sync[2] == self.internal_t
                and sync[1].shape == self.state.shape and np.array_equal(sync[1], self.state, equal_nan=True)):
            eng.set_state(self.state[None], internal_t=self.internal_t)
        out = eng.step_host(act[None])
        self.state[:, 0:dim] = out["pos"][0]
        self.state[:, dim:2 * dim] = out["vel"][0]
        z_states, Ni = self._unpack_obs(out["z"][0], out["Ni"][0])
        self.z_states = z_states
        self.Ni = Ni
        finished = bool(out["fin"][0])
        self.internal_t += 1
        self._synced = (eng, self.state.copy(), self.internal_t)
        return (self.state, z_states, np.array(out["r"][0]), np.int64(out["nc"][0]), finished,
                np.array(out["tr"][0]))

    # ------------------------------------------------------------------ plotting (off-path)
    def show(self, state=None, not_animate=True):
        plt = _plt()
        if not_animate:
            state = self.state
        fig, ax = plt.subplots()
        ax.set_xlim((0, self.grid[0])); ax.set_ylim((0, self.grid[1]))
        ax.set_aspect('equal')
        xF = self.end_points.reshape(self.n_agents, dim)
        for i in range(self.n_agents):
            colour = num_to_rgb(i, max(self.n_agents - 1, 1))
            ax.add_patch(plt.Circle((state[i, 0], state[i, 1]), state[i, 4], color=colour))
            ax.plot(xF[i, 0], xF[i, 1], marker="x", color=colour)
        for o in self.obstacles:
            ax.add_patch(plt.Circle((o[0], o[1]), o[2], color=BLACK))
        if not_animate:                                   # reference drone_env.py:430-433: show, or hand the figure back
            plt.show()
        else:
            return fig

    def animate_basic(self, trajectory, frame_time=0.2, frames=20):
        """reference drone_env.py:436-448: about `frames` evenly spaced states of the trajectory."""
        plt = _plt()
        good_frame = 0
        each_frame = len(trajectory) / frames
        for n_frame, state in enumerate(trajectory):
            if each_frame > 1 and round(good_frame) == n_frame:
                good_frame += each_frame
            else:
                continue
            self.show(state, not_animate=False)
            plt.pause(frame_time)
            plt.close()

    def plot(self, trajectory, episode=None):
        plt = _plt()
        traj = np.array(trajectory)
        fig, ax = plt.subplots()
        ax.set_xlim((0, self.grid[0])); ax.set_ylim((0, self.grid[1])); ax.set_aspect('equal')
        xF = self.end_points.reshape(self.n_agents, dim)
        for i in range(self.n_agents):
            colour = num_to_rgb(i, max(self.n_agents - 1, 1))
            ax.plot(traj[:, i, 0], traj[:, i, 1], color=colour, label=f"Agent {i + 1}")
            ax.plot(traj[0, i, 0], traj[0, i, 1], marker="o", color=colour)
            ax.plot(xF[i, 0], xF[i, 1], marker="x", color=colour)
        ax.set_title("Trajectories" if episode is None else f"Trajectories, episode {episode + 1}")
        ax.legend(); ax.grid(alpha=0.3)
        plt.show()

    def animate(self, trajectory, z_trajectory=None, deltas=None, episode=None, name="test", format="gif"):
        plt = _plt()
        from matplotlib import animation
        import os
        traj = np.array(trajectory)
        fig, ax = plt.subplots()
        ax.set_xlim((0, self.grid[0])); ax.set_ylim((0, self.grid[1])); ax.set_aspect('equal')
        discs = []
        for i in range(self.n_agents):
            colour = num_to_rgb(i, max(self.n_agents - 1, 1))
            disc = plt.Circle((traj[0, i, 0], traj[0, i, 1]), traj[0, i, 4], color=colour)
            ax.add_patch(disc); discs.append(disc)

        def frame(t):
            for i, disc in enumerate(discs):
                disc.center = (traj[t, i, 0], traj[t, i, 1])
            return discs

        anim = animation.FuncAnimation(fig, frame, frames=len(traj), interval=1000 * dt, blit=True)
        os.makedirs("videos", exist_ok=True)
        if format == "gif":
            full_name = os.path.join("videos", name + ".gif")
            anim.save(full_name, writer=animation.PillowWriter(fps=30))
        elif format == "mp4":
            full_name = os.path.join("videos", name + ".mp4")
            anim.save(full_name, writer=animation.FFMpegWriter(fps=30))
        else:
            print(f"format{format} not valid")
            return
        print(f"Animation saved as {full_name}")


# ---------------------------------------------------------------------- controllers (drone_env.py:612-679)
def _device_control(controller, state, env, u_max):
    """Both baseline controllers run on the device (ds_control) for the one environment of `env`."""
    state = np.asarray(state, np.float64)
    # the controllers take the agent radii from state[:, 4] (drone_env.py:632,643,664), not from
    # env.drone_radius: a state whose radius column was edited gets an engine built on those radii
    radii = None if np.array_equal(state[:, 4], env.drone_radius) else state[:, 4]
    eng = env._engine(env.end_points, env.d_safety, env.deltas, radii)
    env._synced = None                                   # the device state is overwritten below
    eng.set_state(np.asarray(state, np.float64)[None], internal_t=env.internal_t)
    act = eng.control(controller, float(u_max))[0].cpu().numpy()
    return [act[i].copy() for i in range(env.n_agents)]


def gradient_control(state, env, u_max=1):
    """Descent direction of the log-barrier cost with global knowledge (drone_env.py:612-653)."""
    return _device_control("gradient", state, env, u_max)


def proportional_control(state, env):
    """Unit-gain P controller with the control norm capped at 1 m/s (drone_env.py:655-679)."""
    return _device_control("proportional", state, env, 1.0)


# ---------------------------------------------------------------------- plotting helpers (drone_env.py:682-741)
def running_average(x, N=50):
    if len(x) >= N:
        y = np.copy(x)
        y[N - 1:] = np.convolve(x, np.ones((N,)) / N, mode='valid')
    else:
        y = np.zeros_like(x)
    return y


def plot_rewards(episode_reward_list, episode_true_reward_list, collision_list, n_ep_running_average=50):
    plt = _plt()
    episodes = list(range(1, len(episode_reward_list) + 1))
    fig, ax = plt.subplots(nrows=1, ncols=2, figsize=(8, 4.5))
    ax[0].plot(episodes, episode_reward_list, label='Global reward', color="cyan", alpha=0.5)
    ax[0].plot(episodes, running_average(episode_true_reward_list, n_ep_running_average),
               label='Avg. global reward', color="blue")
    ax[0].set_xlabel('Episodes'); ax[0].set_ylabel('Total reward')
    ax[0].set_title('Total Reward vs Episodes'); ax[0].legend(); ax[0].grid(alpha=0.3)
    ax[1].plot(episodes, collision_list, label='Collisions per episode', color="cyan", alpha=0.5)
    ax[1].plot(episodes, running_average(collision_list, n_ep_running_average),
               label='Avg. number of collisions per episode', color="blue")
    ax[1].set_xlabel('Episodes'); ax[1].set_ylabel('Total number of collisions')
    ax[1].set_title('Total number of collisions vs Episodes'); ax[1].legend(); ax[1].grid(alpha=0.3)
    plt.show()


def plot_grads(grad_per_episode: np.ndarray, gi_per_episode: np.ndarray):
    plt = _plt()
    fig, ax = plt.subplots(nrows=1, ncols=2, figsize=(16, 9))
    n_agents = np.size(grad_per_episode, 1)
    episodes = list(range(1, len(grad_per_episode) + 1))
    for panel, data, label in ((0, grad_per_episode, 'Score function gradient'),
                               (1, gi_per_episode, 'Approximated gi gradient (max norm = 100)')):
        for i in range(n_agents):
            ax[panel].plot(episodes, data[:, i], label=f"Agent {i + 1}", color=num_to_rgb(i, n_agents - 1))
        ax[panel].set_xlabel('Episodes'); ax[panel].set_ylabel(label)
        ax[panel].legend(); ax[panel].grid(alpha=0.3)
    plt.show()
