from typing import Collection, Hashable, Union

KeysCollection = Union[Collection[Hashable], Hashable]
