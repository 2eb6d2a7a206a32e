"""MiniWorld-Maze-v0 family: recursive-backtracker maze of 3x3 m cells (reference envs/maze.py)."""
from .._gym import spaces, utils
from ..entity import Box
from ..params import DEFAULT_PARAMS
from ..world import MiniWorldEnv
from ._goal import GoalBoxRule


class Maze(GoalBoxRule, MiniWorldEnv, utils.EzPickle):
    def __init__(self, num_rows=8, num_cols=8, room_size=3, max_episode_steps=None, **kwargs):
        self.num_rows, self.num_cols = num_rows, num_cols
        self.room_size = room_size
        self.gap_size = 0.25
        MiniWorldEnv.__init__(self, max_episode_steps=max_episode_steps or num_rows * num_cols * 24, **kwargs)
        utils.EzPickle.__init__(self, num_rows=num_rows, num_cols=num_cols, room_size=room_size,
                                max_episode_steps=max_episode_steps, **kwargs)
        self.action_space = spaces.Discrete(self.actions.move_forward + 1)

    def _gen_world(self):
        pitch = self.room_size + self.gap_size
        grid = []
        for j in range(self.num_rows):
            row = []
            for i in range(self.num_cols):
                min_x, min_z = i * pitch, j * pitch
                row.append(self.add_rect_room(min_x=min_x, max_x=min_x + self.room_size,
                                              min_z=min_z, max_z=min_z + self.room_size,
                                              wall_tex="brick_wall"))
            grid.append(row)
        seen = set()

        def carve(i, j):
            cell = grid[j][i]
            seen.add(cell)
            remaining = [(0, 1), (0, -1), (-1, 0), (1, 0)]
            order = []
            while len(order) < 4:                      # shuffle by repeated choice (4 draws)
                pick = remaining[self.np_random.choice(len(remaining))]
                remaining.remove(pick)
                order.append(pick)
            for dj, di in order:
                ni, nj = i + di, j + dj
                if not (0 <= nj < self.num_rows and 0 <= ni < self.num_cols):
                    continue
                other = grid[nj][ni]
                if other in seen:
                    continue
                if di == 0:
                    self.connect_rooms(cell, other, min_x=cell.min_x, max_x=cell.max_x)
                elif dj == 0:
                    self.connect_rooms(cell, other, min_z=cell.min_z, max_z=cell.max_z)
                carve(ni, nj)

        carve(0, 0)
        self.box = self.place_entity(Box(color="red"))
        self.place_agent()

    def device_program(self, prog):
        """Topology is random per episode: the carving itself is lowered (csrc/maze.cuh), using
        geometry templates that maze_lowering.MazeTemplate cuts out of host-built worlds."""
        prog.maze()
        prog.place(prog.proto(Box(color="red")))
        prog.place_agent()


class MazeS2(Maze):
    def __init__(self, num_rows=2, num_cols=2, **kwargs):
        Maze.__init__(self, num_rows=num_rows, num_cols=num_cols, **kwargs)


class MazeS3(Maze):
    def __init__(self, num_rows=3, num_cols=3, **kwargs):
        Maze.__init__(self, num_rows=num_rows, num_cols=num_cols, **kwargs)


fast_params = DEFAULT_PARAMS.no_random()
fast_params.set("forward_step", 0.7)
fast_params.set("turn_step", 45)


class MazeS3Fast(Maze):
    def __init__(self, num_rows=3, num_cols=3, max_episode_steps=300, params=fast_params,
                 domain_rand=False, **kwargs):
        Maze.__init__(self, num_rows=num_rows, num_cols=num_cols, max_episode_steps=max_episode_steps,
                      params=params, domain_rand=domain_rand, **kwargs)
