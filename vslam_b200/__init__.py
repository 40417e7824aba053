"""Import alias: `import vslam_b200` resolves to the package directory
`vslam-pose-estimation-framework_b200/` (whose name is not a valid Python identifier)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "vslam-pose-estimation-framework_b200"))
